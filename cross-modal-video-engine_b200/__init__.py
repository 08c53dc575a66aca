"""B200-native retrieval scoring engine (drop-in for the LINAS-engine / MultiFusion scoring path)."""
__version__ = "0.1.0"
