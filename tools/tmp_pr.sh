P="--dtype bf16 --graph random --no-ref --points 262144:8:64,65536:8:64,262144:8:128"
for d in 12 8 4; do echo "prune_min_deg=$d"; MG_GAT_PRUNE_MIN_DEG=$d timeout 200 python tools/sweep.py $P --out gpurun_out/pr_$d.md > /dev/null 2>&1; tail -3 gpurun_out/pr_$d.md | cut -d'|' -f3,4,5,7; done
for c in "4099 0 40 64 0 f32" "70000 8 8 64 0 bf16" "5000 0 9 48 1 bf16"; do echo "check $c: $(MG_GAT_PRUNE_MIN_DEG=4 timeout 60 python tools/agg_check.py $c 2>&1 | tail -1)"; done
