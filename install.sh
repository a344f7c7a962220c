#!/usr/bin/env bash
# Build the native libraries in-tree and install the package into the current environment.
#   ./install.sh            build + editable install (no dependency resolution: works offline)
#   ./install.sh --no-build skip the ahead-of-time build (ops compile on first use)
set -euo pipefail
cd "$(dirname "$0")"
if [[ "${1:-}" != "--no-build" ]]; then
    python -c "import __graft_entry__ as g; g.build()"
fi
python -m pip install --no-index --no-build-isolation --no-deps -e .
python -m deepspeed_b200.env_report || true
