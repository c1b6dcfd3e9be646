for v in base minb5 minb6 base minb5 minb6; do
  if [ $v = base ]; then unset PNP_LIB; else export PNP_LIB=$PWD/pnp_svrg_b200/lib/exp/libpnp_$v.so; fi
  python bench.py --sections= --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['us_per_inner_iteration'],2), {k:round(v,2) for k,v in d['kernel_us'].items()})"
done
