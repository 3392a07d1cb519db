"""Import-path alias: the reference's module paths (`python -m oadp.oake.globals ...`) resolve to the
B200-native implementation in `oadp_b200`."""
