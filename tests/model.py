"""Builds small complete SOC models on disk (cloud, dust, dsc, sources, ini) for end-to-end driver tests."""
from soc_b200.synth import write_model  # noqa: F401
