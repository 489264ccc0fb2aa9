#!/bin/sh
# Stage the UNMODIFIED reference model code (pure Python: models/*.py, models/utils/*.py) into
# oracle/_ref/ so that it travels to the GPU box with the repo snapshot.  oracle/_ref/ is git-ignored
# (never committed) and not gpurun-ignored, exactly like a compiled oracle/_ref/*.so would be.
# Test infrastructure only: tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline
# legs may import it (through oracle/reference.py); the product package never does.
#
#     sh oracle/make_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/models" ]; then
    echo "make_ref: $SRC/models not found (fine on the GPU box: oracle/_ref is prebuilt)" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/models/utils"
cp "$SRC"/models/*.py "$DST/models/"
cp "$SRC"/models/utils/*.py "$DST/models/utils/"
( cd "$SRC" && sha256sum models/*.py models/utils/*.py ) > "$DST/SHA256SUMS"
echo "make_ref: staged $(ls "$DST"/models/*.py "$DST"/models/utils/*.py | wc -l) files into $DST"
