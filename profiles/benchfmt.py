import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): 
        if l: print(l)
        continue
    d=json.loads(l)
    r=d.get("roofline") or {}
    print("value %.0f conv/s  ms/step %.2f  e2e %.0f  kernel_ms %s  conv_launch_avg_ms %.4f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_ms"), r.get("launch_ms_avg",0), d.get("gpu_launches")))
