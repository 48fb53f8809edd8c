# round 2, call 21 (needs the BSG_S2_PROMO switch of commit "profiles: N=1 bench line and per-layer table of the current build; L2 promotion switch"): L2 promotion of the stride-2 parity-view tensor maps
cd "$GRAFT_REPO_ROOT"
for p in 256 128 64 0; do
  BSG_S2_PROMO=$p timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers21_promo$p.log 2>&1; echo "promo=$p rc=$?"; grep "back-to-back\|conv3 s2" gpurun_out/r02_layers21_promo$p.log | cut -c1-100
done
