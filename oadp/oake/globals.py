"""`python -m oadp.oake.globals NAME CONFIG` -- same entry point as the reference."""
from oadp_b200.oake.globals import *  # noqa: F401,F403
from oadp_b200.oake.globals import Validator

if __name__ == '__main__':
    Validator.main()
