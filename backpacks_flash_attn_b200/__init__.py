"""backpacks_flash_attn_b200: the Backpack forward hot path (fused attention + sense-mix) for B200."""
__version__ = "0.1.0"
