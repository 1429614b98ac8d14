mkdir -p gpurun_out
T=${TAG:-r2i}
timeout 1200 python -m pytest tests -m gpu -x -q -k "wide or dem or encoder or dynamic or rescal" > gpurun_out/${T}_pytest_wide.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_wide.log
tail -3 gpurun_out/${T}_pytest_wide.log
timeout 600 python benchmarks/wide_d5.py 5 8 > gpurun_out/${T}_wide_d5_bfd.jsonl 2>&1; grep rep gpurun_out/${T}_wide_d5_bfd.jsonl | cut -c1-300
TQEC_WIDE_BF_STAGED=1 timeout 600 python benchmarks/wide_d5.py 5 8 2>&1 | grep rep | cut -c1-300
TQEC_WIDE_BF_CTAS=3 timeout 600 python benchmarks/wide_d5.py 5 8 2>&1 | grep rep | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_wide_bfd -s 60 -c 1 -o gpurun_out/${T}_wide_bfd -f python benchmarks/wide_d5.py 5 1 > gpurun_out/${T}_ncu_wide_bfd.log 2>&1
tail -2 gpurun_out/${T}_ncu_wide_bfd.log
