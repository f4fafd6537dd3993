# Segmentation forward chunk size (tiles per chunk of one forward call) against the C2 step: device-resident value only.
for c in 148 192 384 148 384; do
  BDRC_SEG_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $c: value %.1f pages/s  %.3f ms/step  e2e %.1f  conv frac %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))"
done
