"""Runs pytest against an alternative build of the device library: tools/pytest_with_lib.py <lib file name> [pytest args]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuking_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), sys.argv[1])
import pytest
raise SystemExit(pytest.main(sys.argv[2:]))
