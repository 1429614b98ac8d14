mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tnmmap or library_compile or sum_product or sumprod or marginal" > gpurun_out/r3d_pytest.log 2>&1; tail -3 gpurun_out/r3d_pytest.log
python benchmarks/tnmmap_quick.py 2>&1 | tail -3 | cut -c1-220
python - <<'PY'
import time, sys, os
sys.path.insert(0, os.getcwd())
import tensorqec.jl_b200 as tq
t = tq.CSSTannerGraph(tq.SurfaceCode(9, 9)); em = tq.iid_error(0.05, t)
t0 = time.time(); ct = tq.compile(tq.TNMMAP(), t, em); print("compile TNMMAP d=9 s", round(time.time() - t0, 2))
t0 = time.time(); ct = tq.compile(tq.TNMAP(), t, em); print("compile TNMAP d=9 s", round(time.time() - t0, 2))
PY
