"""bioframe_b200: B200-native engine behind bioframe's interval-operation API."""
__version__ = "0.1.0"
