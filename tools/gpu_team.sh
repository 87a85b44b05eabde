#!/bin/bash
# One GPU iteration on the team schedule: parity tests, bench sweep over team geometries, phase clocks.
# Usage: tools/gpu_team.sh <tag> [test-filter]
tag=$1; filt=${2:-"team or full_size or reference_trees"}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -k "$filt" 2>&1 | tail -8 > gpurun_out/test_$tag.log; tail -4 gpurun_out/test_$tag.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pl=d["plan"]
    print("%-22s kern_ms %.3f value %.3e e2e %.3e"%(sys.argv[2],d["roofline"]["kernel_ms"],d["value"],d["e2e"]["value"]), pl.get("schedule"), "G",pl.get("lanes_per_tree"),"tpt",pl.get("trees_per_team"),"teams",pl.get("teams_per_cta"),"nw",pl.get("net_warps_per_team"),"blk",pl["block"],"smem",pl["smem_bytes"])
except Exception as e:
    print(sys.argv[2],"FAILED",e, open(sys.argv[1]).read()[-300:], open(sys.argv[1][:-5]+".err").read()[-300:])
PY
}
run() { name=$1; shift; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/bench_${tag}_$name.json 2>gpurun_out/bench_${tag}_$name.err; show gpurun_out/bench_${tag}_$name.json "$name"; }
clk() { name=$1; shift; timeout 300 python bench.py --phase-clocks --no-cpu "$@" > gpurun_out/clk_${tag}_$name.json 2>&1; python - gpurun_out/clk_${tag}_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["cycles_per_simulation_mean"]
    print("clk %-14s"%sys.argv[2]," ".join("%s=%.0f"%(k.replace("tree.","T.").replace("net.","N."),v) for k,v in c.items()))
except Exception as e:
    print("clk",sys.argv[2],"FAILED",e,open(sys.argv[1]).read()[-300:])
PY
}
source tools/gpu_team_cases.sh
