"""B200-native voxel_plus hot path (IEKF measurement update + voxel-map update)."""
