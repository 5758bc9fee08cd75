import json,sys
lines=[l for l in open(sys.argv[1]) if l.startswith("{")]
d=json.loads(lines[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],2),d["clocks"], "frac", round(d["model_frac_of_bf16_sustained"],3))
for k in ("bags_identical","bags_digest","collective","config3_uniform_40x50","weak_16_regions_per_rank"):
    if k in d: print(" ",k,d[k])
print("  e2e",{k:v for k,v in d["e2e"].items() if k!="host_pool"})
c=d["config"]; print("  regions/rank",c.get("regions_per_rank"),"cut",c.get("slides_cut_by_a_rank_boundary"),"slides",c.get("slides_per_step"))
for k,v in d["kernels"].items():
    if v["share"]>0.004: print(" ",k, round(v.get("ms_per_region",v.get("ms_per_step",0)),3), round(v["share"],3), round(v.get("tflops",0),0), "us/launch",round(v["avg_launch_us"],1))
for k in ("vit256_config2","clam_config4","cpu_baseline"):
    if k in d: print(" ",k,json.dumps(d[k])[:900])
