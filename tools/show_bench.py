import json,sys
lines=[l for l in open(sys.argv[1]) if l.startswith("{")]
d=json.loads(lines[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],2),d["clocks"], "frac", round(d["model_frac_of_bf16_sustained"],3))
for k,v in d["kernels"].items():
    if v["share"]>0.004: print(" ",k, round(v["ms_per_step"],2), round(v["share"],3), round(v.get("tflops",0),0), round(v.get("gbs",0),0))
