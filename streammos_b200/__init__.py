"""streammos_b200 — B200-native (sm_100a) implementation of StreamMOS's per-scan hot path behind the
reference's own operator API. See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
