"""Ops tooling: bus publisher, allow-list updater, feature generator, resource monitors."""
