"""pytest plugin (opt-in): replaces the library binding with the dry-run stand-in, see mocklib.py.

    PYTHONPATH=tests/dryrun python -m pytest -p mockplugin tests/test_zz_* -m gpu -p no:cacheprovider
"""


def pytest_configure(config):
    import mocklib

    mocklib.install()
