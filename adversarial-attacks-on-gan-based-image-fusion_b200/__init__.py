"""sfattack-b200: B200-native attack hot path."""
__version__ = "0.1.0"
