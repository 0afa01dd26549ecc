"""Bare-name shim: `import basic` after putting disentangledcolorization_b200/compat on sys.path (the reference's
main/_init_paths.py:10-13 import style) resolves to the B200 drop-in module."""
from disentangledcolorization_b200.basic import *  # noqa: F401,F403
