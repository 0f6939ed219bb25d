"""`python main.py ...` -- the reference's command line on the native engine."""
from dorpatch_b200.main import cli, main, parser  # noqa: F401

if __name__ == '__main__':
    cli()
