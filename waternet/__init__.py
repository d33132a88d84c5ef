"""Reference-compatible import paths (``waternet.net``, ``waternet.data``,
``waternet.training_utils``) resolving to the B200-native implementation."""
