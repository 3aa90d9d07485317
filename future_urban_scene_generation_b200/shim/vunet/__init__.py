"""`vunet` as the reference's callers import it (run_test.py:20), served by the B200 path."""
