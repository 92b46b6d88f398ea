#!/bin/bash
python - <<'PY'
import time, sys, os, ctypes
t0=time.perf_counter()
rt=ctypes.CDLL("libcudart.so.12")
rt.cudaFree(0)
t1=time.perf_counter(); print("cudaFree(0) = context creation: %.3f s"%(t1-t0))
sys.path.insert(0,'.')
from loom_b200 import _abi, Engine, synth
cuda=_abi.load_cuda(); t2=time.perf_counter(); print("dlopen libloomb200: %.3f s"%(t2-t1))
e=Engine(cuda); t3=time.perf_counter(); print("lb_create: %.3f s"%(t3-t2))
m,r,_=synth.generate(2000,[("bb",5),("dd256",5),("gp",5),("nich",5)],seed=1); t4=time.perf_counter()
e.set_model(m); t5=time.perf_counter(); print("lb_set_model: %.3f s"%(t5-t4))
e.set_rows(r); t6=time.perf_counter(); print("lb_set_rows: %.3f s"%(t6-t5))
from loom_b200 import sweep_actions
e.cat_sweep(sweep_actions(2000),seed=1); t7=time.perf_counter(); print("first sweep (module load of the sweep kernels): %.3f s"%(t7-t6))
e.cat_sweep(sweep_actions(2000,add=True,remove=True),seed=2); t8=time.perf_counter(); print("second sweep: %.3f s"%(t8-t7))
PY
nvidia-smi --query-gpu=persistence_mode --format=csv,noheader
