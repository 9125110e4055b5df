// reference-native descriptor path (filled in below)
