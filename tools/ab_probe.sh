#!/bin/bash
# A/B of development builds of the fused kernel: tools/ab_probe.sh <n> <m> lib1 lib2 ...   (libs live in build_variants/)
n=$1; m=$2; shift 2
for lib in "$@"; do
  echo "== $lib"
  NK_LIB_PATH=$PWD/build_variants/$lib timeout -s KILL 300 python tools/perf_probe.py $n $m 192 6 ${CHUNK:-512} 2 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print(('   check ok=%s max_rel_err=%.1e' % (d['ok'], d['max_rel_err'])) if 'check' in d else ('   %.1f ms  %.0f samples/s  frac %.4f' % (d['ms'], d['samples_per_s'], d['frac_of_37_1'])))
    except Exception: print(l.strip())
"
done
