#!/bin/bash
# Local wrapper: rebuild every native artefact (so that the .so files that travel are current), then gpurun.
#   scripts/grun.sh [--gpus N] <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
GP=()
if [ "$1" == "--gpus" ]; then GP=(--gpus "$2"); shift 2; fi
python -c "import __graft_entry__ as g; g.build()"
make -s -C host 2>/dev/null || true
exec /usr/local/graft/bin/gpurun "${GP[@]}" --timeout "$1" -- "$2"
