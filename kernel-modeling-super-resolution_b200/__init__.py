"""kmsr_b200 -- B200-native LR/HR training-pair synthesis (drop-in for the reference hot path)."""
__version__ = "0.1.0"
