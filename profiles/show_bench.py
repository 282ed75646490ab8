import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in ("configs4_n1","configs2_n1","configs3_n1"):
    v=d.get(k); print(k, {kk: v.get(kk) for kk in ("value","ms_per_step","error")} if v else None, (v or {}).get("e2e",{}).get("value"))
