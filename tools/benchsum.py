import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    B=d["config"]["blocks_per_step"]
    print("B=%d value %.0f MS/s  ms/step %.4f  us/block %.2f  e2e %.0f"%(B,d["value"],d["ms_per_step"],1e3*d["ms_per_step"]/B,d["e2e"]["value"]))
    for k,v in d["kernels"].items(): print("   %-10s %.2f us/block  alg %.0f GB/s"%(k, 1e3*v["avg_ms"]/B, v["alg_gbs"]))
    print("   pipeline frac %.3f wall %.3f"%(d["roofline_pipeline"]["frac"], d["roofline_pipeline"]["frac_wall"]))
