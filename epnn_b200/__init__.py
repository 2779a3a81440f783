"""epnn_b200: B200-native EPNN charge inference behind the reference's entry points."""
__version__ = "0.1.0"
