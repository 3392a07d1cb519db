"""`python -m oadp.oake.objects NAME CONFIG` -- same entry point as the reference."""
from oadp_b200.oake.objects import *  # noqa: F401,F403
from oadp_b200.oake.objects import Validator

if __name__ == '__main__':
    Validator.main()
