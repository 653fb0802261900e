"""fastposecnn_b200 -- B200-native pose-recovery path of FastPoseCNN (placeholder, filled below)."""
