#!/bin/bash
# Pin the oracle (and the goldens it produced) against the REAL reference, wherever a Go toolchain exists.
#
#   tests/golden/verify_with_go.sh [/path/to/reference/checkout]     (default /root/reference)
#
# 1. copies the reference checkout to a scratch directory (it is read-only here) and drops go/cmd/dump/main.go
#    into it as cmd/dump -- the dumper imports the reference's own pkg/fluid, nothing of this repository;
# 2. runs it for every case of tests/golden/make_golden.py that uses the reference's solver
#    (the *_redblack case is this repository's own ordering and has no Go counterpart);
# 3. compares the raw float32 dumps with the committed .npz bit for bit (tests/golden/compare_go_dump.py).
# Exit 0 = every field of every case identical; 3 = no Go toolchain (nothing verified); 1 = mismatch.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference}
if ! command -v go >/dev/null 2>&1; then
    echo "verify_with_go: no Go toolchain on PATH (go version failed): goldens stay pinned by the oracle only" >&2
    exit 3
fi
go version
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/ref"
chmod -R u+w "$WORK/ref"
mkdir -p "$WORK/ref/cmd/dump"
cp "$REPO/go/cmd/dump/main.go" "$WORK/ref/cmd/dump/main.go"
cd "$WORK/ref"
# pkg/fluid needs only the standard library; -mod=mod keeps the build from touching the ebiten dependency of main/
export GOFLAGS=-mod=mod GOAMD64=v1
go build -o "$WORK/dump" ./cmd/dump
# the reference's own tests first: the oracle restates their assertions (tests/test_reference_suite.py)
go test ./pkg/fluid/ 2>&1 | tail -3 || true
run() {   # name preset w h steps [flags...]
    local name=$1 preset=$2 w=$3 h=$4 steps=$5; shift 5
    "$WORK/dump" -preset "$preset" -w "$w" -h "$h" -steps "$steps" -out "$WORK/out/$name" "$@"
}
run jet_130x66 jet 130 66 1,2,10,100
run cavity_96x96_bfecc cavity 96 96 1,10,100 -bfecc
run karman_160x80_bfecc_conf karman 160 80 1,10,100 -bfecc -confinement 0.1
run jet_300x251_default jet 300 251 100
python "$HERE/compare_go_dump.py" "$WORK/out" "$HERE"
