import sys, os
sys.path.insert(0, "/root/repo")
from cuking_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import bench
bench.main()
