timeout 200 python tools/bench_conv.py --mask 1 --only fwd,dgrad --layers 2,3,4,5,6,14 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(r['layer'], r['op'], r['ms'], r['TFLOPs'], r['min_GBps'])
"
