#!/bin/bash
# Sweep of the two-pass GroupNorm grid geometry (norm.cu gn_geometry): bytes per CTA x cap of chunks per sample, on the
# VAE shapes of tools/bench_groupnorm.py. The cap is clamped to the library's kGNMaxChunks (raise it there to sweep beyond).
# Result of the round-2 sweep (kGNMaxChunks = 1024 at the time): profiles/r02b_gn_geometry_sweep.txt.
for kb in 32 64 128; do for mc in 128 222 256 384 512 1024; do
echo "== chunk_kb $kb max_chunks $mc"
SDEO_GN_CHUNK_KB=$kb SDEO_GN_MAX_CHUNKS=$mc timeout 100 python tools/bench_groupnorm.py --only vae --no-ref --no-partner --iters 10 2>&1 | grep vae | cut -c1-110
done; done
