"""Host-side mirror of the reference's `oadp.oake` package (same class names, CLI, configs and
output files; oadp/oake/{base,globals,blocks,objects}.py) on top of the sm_100a pipeline."""
from . import base, blocks, globals, objects  # noqa: F401,A004
