for k in 4 6 8; do echo "lanes $k"; NMR_VIEW_LANES=$k timeout 300 python tools/run_configs.py --configs 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('seconds','views_per_s','mrays_per_s_e2e') if k in d})"; done
