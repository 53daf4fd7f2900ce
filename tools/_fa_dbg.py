import os, sys
sys.path.insert(0, '/root/repo')
import importlib.util
spec = importlib.util.spec_from_file_location('atc', '/root/repo/tools/attn_tc_check.py')
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
m.timeit(1, 32, 96, 8192, True, 1, iters=1)
os.environ['P3_FA_DBG'] = '1'
m.timeit(1, 32, 96, 8192, True, 1, iters=1)
