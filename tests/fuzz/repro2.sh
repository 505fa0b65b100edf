r() { echo -n "$* => "; timeout 120 python tests/fuzz/repro_case.py "$@" 2>&1 | tail -1; }
for D in 96 144 160 176 192 208 224 240; do
for mode in 0 1 2; do
r 400 32 2 tex 0,$D,5,200,800,1,63,12,0,0,$mode
done; done
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 fused_sweep=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 cost_fusion=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 path_hh=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 path_hh=1
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 path_hh=0 fused_sweep=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 h_ws=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 h_ws=0 fused_sweep=0
r 400 32 2 tex 0,192,5,200,800,1,63,12,0,0,0 h_ws=0 path_hh=0 fused_sweep=0 h_legacy=1
