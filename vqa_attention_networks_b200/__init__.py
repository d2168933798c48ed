"""B200-native MFB/MFH fusion and co-attention hot path."""
