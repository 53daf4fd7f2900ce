import os, sys
sys.path.insert(0, '/root/repo')
import torch
sys.argv = ['x']
import importlib.util
spec = importlib.util.spec_from_file_location('atc', '/root/repo/tools/attn_tc_check.py')
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(m.timeit(1, 32, 96, 8192, True, 1, iters=1))
