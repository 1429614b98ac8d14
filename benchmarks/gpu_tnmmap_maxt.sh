# sum-product k_sweep (TNMMAP, CSS) over the register-budget variants: 512 / 640 / 768 threads per CTA
for mt in 512 640 768; do echo "== MAXT $mt"; TQEC_SWEEP_MAXT=$mt python benchmarks/tnmmap_quick.py 2>&1 | grep case | cut -c1-160; done
