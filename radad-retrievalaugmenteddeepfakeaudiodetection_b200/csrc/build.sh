#!/usr/bin/env bash
# Build libradad_flat.so (C ABI of include/radad_flat.h) for sm_100a, in-tree (see Makefile; objects in csrc/build/).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
make -C "${HERE}" -j"$(nproc)" "$@"
