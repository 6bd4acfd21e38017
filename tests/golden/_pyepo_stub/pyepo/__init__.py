"""Three-symbol stand-in for PyEPO (absent from this image), used ONLY by make_golden.py so the
unmodified reference src/cave.py can be imported.  Symbols inferred from the reference's call
sites: src/cave.py:17-18, 53, 62-67, 73, 126, 195, 201; test/test_func.py:21-29."""
from enum import Enum


class EPO(Enum):
    MINIMIZE = 1
    MAXIMIZE = -1
