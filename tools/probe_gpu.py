"""Instruction-form throughput probes (ops/s over all SMs) -> gpurun_out/pipe_probe.json"""
import json, sys
sys.path.insert(0, ".")
import bpperm_b200
be = bpperm_b200.Backend(0)
names = {0: "IMAD.WIDE.U32 plain (64-bit addend)", 1: "IMAD.WIDE.U32 carry-chained (.X / carry-out)",
         2: "IMAD 32-bit lo", 3: "IADD3.X carry chain", 4: "fe_mul chain (x72 IMAD.WIDE)",
         5: "ge_madd chain (x504 IMAD.WIDE)"}
res = {}
for mode, name in names.items():
    it = 4096 if mode < 4 else 1024
    vals = [be.pipe_probe(mode, it) for _ in range(3)]
    res[name] = max(vals)
    print(f"{name:50s} {max(vals)/1e12:8.3f} Tops/s  = {max(vals)/148/1.965e9:6.2f} lanes/clk/SM @1965MHz")
json.dump(res, open("gpurun_out/pipe_probe.json", "w"), indent=1)
