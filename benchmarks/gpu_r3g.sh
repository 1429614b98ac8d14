python -m pytest tests -m gpu -x -q -k "byte or api or error_behaviour or cabi or small_codes or tnmmap_golden" 2>&1 | tail -2
for c in 8 6 8; do echo "copy threads $c"; TQEC_COPY_THREADS=$c python benchmarks/api_profile.py 2>&1 | grep -E "total|decode_map_bits"; done
