"""`from criterion import build_criterion` (reference main.py:22)."""
from coda_neurips2023_b200.criterion import Matcher, SetCriterion, build_criterion  # noqa: F401
