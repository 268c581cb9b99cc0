"""jda_b200: B200-native JDA detect path."""
