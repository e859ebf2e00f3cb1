#!/bin/bash
# Same-box A/B: exports the tree of a git ref into .ab_prev/ (git-ignored, travels with gpurun) and builds its
# libtrb.so, so that `bash profiles/ab_run.sh` can time both versions back to back on ONE GPU box (box-to-box
# spread of the 64-view step is ~8%, larger than most single optimisations).
#   bash profiles/ab_setup.sh <git-ref>
set -e
ref=${1:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
rm -rf "$root/.ab_prev"; mkdir -p "$root/.ab_prev"
git -C "$root" archive "$ref" | tar -x -C "$root/.ab_prev"
cp "$root/MEASURED_PEAKS.json" "$root/.ab_prev/" 2>/dev/null || true
(cd "$root/.ab_prev" && python -m torch_renderer_b200.build > /dev/null && echo "built $ref in .ab_prev")
