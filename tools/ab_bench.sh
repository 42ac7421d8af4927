#!/bin/bash
# A/B run of the library builds under ab_libs/ (tools/ab_build.py) on ONE GPU box: parity first, then the bench, each build alone.
#   gpurun --timeout 1500 -- 'bash tools/ab_bench.sh'      results: gpurun_out/ab_<name>.json / .log
# The split encoder is also run switched off (GZB_AR_SPLIT_MIN=off) inside its own build: same binary, one variable.
set -u
mkdir -p gpurun_out
STEPS=${STEPS:-3}; WARMUP=${WARMUP:-3}
for d in ab_libs/*/; do
    name=$(basename "$d"); lib="$PWD/$d/libgzb200.so"
    [ -f "$lib" ] || continue
    echo "== $name ($(cat "$d/REV" 2>/dev/null))"
    GZB200_LIB="$lib" timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "gpurun_out/ab_${name}.log" 2>&1
    echo "   pytest -m gpu: exit $? ($(tail -1 gpurun_out/ab_${name}.log))"
    GZB200_LIB="$lib" timeout 600 python bench.py --steps "$STEPS" --warmup "$WARMUP" > "gpurun_out/ab_${name}.json" 2>> "gpurun_out/ab_${name}.log"
    echo "   bench: $(python -c "import json,sys; d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]); print('value', d['value'], 'zip', d.get('zip_GBps'), 'piz', d.get('piz_GBps'), '| e2e', d['e2e']['value'], '| ms/step', d['ms_per_step'], '| VBlocks/step', d['config'].get('vblocks_per_gpu_per_step'))" 2>/dev/null || echo failed)"
    if [[ "$name" == *split* ]]; then
        GZB_AR_SPLIT_MIN=off GZB200_LIB="$lib" timeout 600 python bench.py --steps "$STEPS" --warmup "$WARMUP" > "gpurun_out/ab_${name}_off.json" 2>> "gpurun_out/ab_${name}.log"
        echo "   bench with GZB_AR_SPLIT_MIN=off: $(python -c "import json; d=json.loads(open('gpurun_out/ab_${name}_off.json').read().strip().splitlines()[-1]); print('value', d['value'], '| e2e', d['e2e']['value'], '| VBlocks/step', d['config'].get('vblocks_per_gpu_per_step'))" 2>/dev/null || echo failed)"
    fi
done
# the vectorised host driver (Python side): its tree under ab_libs/_vhd_tree, main's library
if [ -d ab_libs/_vhd_tree ]; then
    lib="$PWD/ab_libs/main/libgzb200.so"
    ( cd ab_libs/_vhd_tree && GZB200_LIB="$lib" timeout 600 python bench.py --steps "$STEPS" --warmup "$WARMUP" ) > gpurun_out/ab_vhd.json 2> gpurun_out/ab_vhd.log
    ( cd ab_libs/_vhd_tree && GZB200_LIB="$lib" timeout 600 python bench.py --pipelines --steps "$STEPS" --warmup "$WARMUP" ) > gpurun_out/ab_vhd_pipelines.json 2>> gpurun_out/ab_vhd.log
    for f in ab_vhd ab_vhd_pipelines; do echo "   $f: $(python -c "import json; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('value', d['value'], '| e2e', d['e2e']['value'], '| zip', d.get('zip_GBps'), 'piz', d.get('piz_GBps'))" 2>/dev/null || echo failed)"; done
fi
