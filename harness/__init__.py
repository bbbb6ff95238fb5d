"""Teacher architectures hosted around the hot path for measurement and step-level parity tests (not product code)."""
