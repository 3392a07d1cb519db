from oadp_b200.oake import base, blocks, globals, objects  # noqa: F401,A004
