# round 2, call 26 (needs the BSG_L2_FETCH switch, removed again afterwards): DRAM -> L2 fetch granularity hint
# does the 2x DRAM over-read of the 32-channel stride-2 conv (its input is one half of every 128-byte line of the concat
# buffer) come from whole-line fetches?
cd "$GRAFT_REPO_ROOT"
for g in 128 64 32; do
  BSG_L2_FETCH=$g timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers26_fetch$g.log 2>&1; echo "fetch=$g rc=$?"; grep "back-to-back\|step   2 \|step   1 conv3 s1 32\|step  26 conv3 s1 32" gpurun_out/r02_layers26_fetch$g.log | cut -c1-100
done
