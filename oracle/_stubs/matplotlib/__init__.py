"""Empty stand-in for matplotlib (imported by the reference's utils.py:1, examples.py:3;
not on the numeric path).  TEST INFRASTRUCTURE ONLY, used by oracle/make_golden.py."""
