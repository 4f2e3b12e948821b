"""B200-native DM-VIO photometric hot path (package root)."""
