#!/bin/bash
# Build libplanet_gpu.so variants with extra nvcc defines (K2 tuning experiments) into
# planet_b200/variants/<name>.so; `run` times each on the C2 batch (on the GPU box) by copying
# it over the in-tree library.   tools/k2_variants.sh build NAME "-DFOO=1" ... | run
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
    name=$2; shift 2
    PLANET_NVCC_EXTRA="$*" python planet_b200/build.py --force > /dev/null
    cp planet_b200/libplanet_gpu.so planet_b200/variants/$name.so
    echo "built $name ($*)"
elif [ "$1" = run ]; then
    cp planet_b200/libplanet_gpu.so /tmp/libplanet_gpu.keep
    for so in planet_b200/variants/*.so; do
        cp $so planet_b200/libplanet_gpu.so; touch planet_b200/libplanet_gpu.so
        echo "== $(basename $so .so)"
        for T in ${K2_THREADS_LIST:-768}; do PLANET_K2_THREADS=$T K2_KIND=fbm python tools/k2_sweep.py child | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('ms','gvert_s','tflops','checksum')}), d['env'].get('PLANET_K2_THREADS'))"; done
    done
    cp /tmp/libplanet_gpu.keep planet_b200/libplanet_gpu.so; touch planet_b200/libplanet_gpu.so
fi
