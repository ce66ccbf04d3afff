"""vmad_b200: B200-native particle-mesh backend behind vmad's operator contract."""
