# Butterfly executor (k_wide_bf): parity tests of every wide / DEM plan, then d = 5 x 5 rounds with and without it.
mkdir -p gpurun_out
T=${TAG:-r2e}
timeout 1200 python -m pytest tests -m gpu -x -q -k "wide or dem or encoder or dynamic or rescal" > gpurun_out/${T}_pytest_wide.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_wide.log
tail -6 gpurun_out/${T}_pytest_wide.log
timeout 600 python benchmarks/wide_d5.py 5 8 > gpurun_out/${T}_wide_d5_bf.jsonl 2>&1; cut -c1-300 gpurun_out/${T}_wide_d5_bf.jsonl
TQEC_WIDE_NO_BF=1 timeout 600 python benchmarks/wide_d5.py 5 8 > gpurun_out/${T}_wide_d5_nobf.jsonl 2>&1; cut -c1-300 gpurun_out/${T}_wide_d5_nobf.jsonl
for c in ${CTAS:-3 2}; do echo "== bf ctas/sm $c"; TQEC_WIDE_BF_CTAS=$c timeout 600 python benchmarks/wide_d5.py 5 8 2>&1 | grep rep | cut -c1-200; done
timeout 600 python benchmarks/dem_wide_compare.py 2>&1 | tail -8 | cut -c1-300
