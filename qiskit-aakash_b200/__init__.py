"""B200-native dm_simulator hot path."""
