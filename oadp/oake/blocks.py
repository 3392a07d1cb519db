"""`python -m oadp.oake.blocks NAME CONFIG` -- same entry point as the reference."""
from oadp_b200.oake.blocks import *  # noqa: F401,F403
from oadp_b200.oake.blocks import Validator

if __name__ == '__main__':
    Validator.main()
