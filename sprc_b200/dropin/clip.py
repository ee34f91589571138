"""`import clip` succeeds (src/blip_validate.py:8 imports OpenAI CLIP but never uses it)."""
