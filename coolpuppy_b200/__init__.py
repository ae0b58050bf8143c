"""B200-native pile-up engine behind the coolpuppy API."""
__version__ = "0.1.0"
