# sum-product sweep: head bits 10 (shipped) / 12 / 14 at d = 9 and d = 7
for hb in 10 12 14 10 12; do
  echo "== sp head bits $hb"
  TQEC_HEAD_BITS_SP=$hb python benchmarks/tnmmap_quick.py 9 400000 2>&1 | tail -1 | cut -c1-200
done
TQEC_HEAD_BITS_SP=12 python benchmarks/tnmmap_quick.py 7 1000000 2>&1 | tail -1 | cut -c1-200
TQEC_HEAD_BITS_SP=10 python benchmarks/tnmmap_quick.py 7 1000000 2>&1 | tail -1 | cut -c1-200
