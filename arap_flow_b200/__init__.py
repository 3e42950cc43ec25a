"""arap_flow_b200 -- B200-native ARAP flow generator (hot path of lhoangan/arap_flow)."""
__version__ = "0.1.0"
