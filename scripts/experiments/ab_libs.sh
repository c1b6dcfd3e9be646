# A/B of library builds through the default bench command: bash scripts/experiments/ab_libs.sh base <variant> base <variant> ...
# (variants are pnp_svrg_b200/lib/exp/libpnp_<variant>.so, built with PNP_LIB_OUT=... PNP_NVCC_EXTRA=... python -m pnp_svrg_b200.build --force)
for v in "$@"; do
  if [ $v = base ]; then unset PNP_LIB; else export PNP_LIB=$PWD/pnp_svrg_b200/lib/exp/libpnp_$v.so; fi
  python bench.py --sections= --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['us_per_inner_iteration'],2), {k:round(v,2) for k,v in d['kernel_us'].items()}, d['psnr_first_last'])"
done
