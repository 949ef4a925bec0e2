#!/bin/bash
for m in 1 0; do echo "SDF_LIF_BWD_PRELOAD=$m"; SDF_LIF_BWD_PRELOAD=$m python tools/microbench.py 2>&1 | grep "lif_bwd+bn T=10" | cut -c1-200; done
