"""Module-level profiling switches, kept for import compatibility with ``pyDNMFk/config.py:1-5``."""
time = {}
flag = 0


def init(arg):
    """Reset the shared timing state (reference: config.py:1-5)."""
    global time, flag
    time = {}
    flag = 0
