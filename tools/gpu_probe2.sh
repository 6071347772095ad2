for c in -1 100 50; do echo "carveout $c"; VDL2GPU_CARVEOUT=$c python tools/probe_overlap.py 2>&1 | grep -E "overlap|serial host"; done
