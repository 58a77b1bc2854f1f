"""Test-only stand-in for the parts of `timm` that the reference's encoder imports (utils/mim_vit.py:6-8):
PatchEmbed, Block, AttentionPoolLatent and optim_factory.param_groups_weight_decay.  timm is not installed in this
image; the search path only consumes the latents the encoder emits, and both sides of every
comparison run the SAME encoder, so the shim only has to be a standard ViT of the right shapes.  It is put on sys.path
by tests/ref_encoder.py and by oracle/make_golden.py, never by the product."""
__version__ = "0.0-shim"
