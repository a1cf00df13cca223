"""Pretty-print the JSON line of a bench log:  python tools/show_bench.py gpurun_out/b1.log [key-prefix ...]"""
import json, sys
l = json.loads([x for x in open(sys.argv[1]).read().strip().splitlines() if x.startswith("{")][-1])
pref = sys.argv[2:]
def show(d, ind=0, path=""):
    for k, v in d.items():
        p = path + k
        if isinstance(v, dict):
            if not pref or any(p.startswith(q) or q.startswith(p) for q in pref):
                print(" " * ind + k + ":"); show(v, ind + 2, p + ".")
        elif not pref or any(p.startswith(q) for q in pref):
            print(" " * ind + f"{k}: {str(v)[:150]}")
show(l)
