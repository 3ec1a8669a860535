#!/usr/bin/env bash
# Build an A/B variant of libdnmf.so:  tools/build_variant.sh NAME "-DKL_S1=0 -DDNMF_WAIT_HINT=0"
# Only the tcgen05 translation units are recompiled (into csrc/build_NAME/); the other objects come from the main build.
# The result is pydnmfk_b200/libdnmf_NAME.so, selected at run time with DNMF_LIB_PATH (tools/kl_lab.py does that).
set -euo pipefail
NAME="$1"; DEFS="${2:-}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../pydnmfk_b200/csrc" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
bash "${HERE}/build.sh" > /dev/null
mkdir -p "${HERE}/build_${NAME}"
pids=()
for f in dnmf_tc dnmf_tc_kl dnmf_tc_kl64; do
  "${NVCC}" "${FLAGS[@]}" ${DEFS} -c -o "${HERE}/build_${NAME}/${f}.o" "${HERE}/${f}.cu" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "${p}"; done
objs=()
for o in "${HERE}"/build/*.o; do
  b="$(basename "${o}")"
  if [[ -f "${HERE}/build_${NAME}/${b}" ]]; then objs+=("${HERE}/build_${NAME}/${b}"); else objs+=("${o}"); fi
done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${HERE}/../libdnmf_${NAME}.so" "${objs[@]}" -lcudart_static -ldl -lpthread -lrt
echo "built libdnmf_${NAME}.so (${DEFS})"
