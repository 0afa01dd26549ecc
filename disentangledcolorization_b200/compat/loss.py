"""bare-name `loss` module (main/_init_paths.py:10-13 puts models/ on sys.path)."""
from disentangledcolorization_b200.loss import *  # noqa: F401,F403
from disentangledcolorization_b200.loss import AnchorColorProbLoss, SPixelLoss  # noqa: F401
