#!/bin/bash
# Produce tests/golden/ref_*.npz from the real reference under lavapipe.  See README.md (untested in the build image).
set -euo pipefail
REF=${1:?path to a num3ric/sol-rs checkout}
REPO=${2:?path to this repository}
HERE="$(cd "$(dirname "$0")" && pwd)"
WORK=${TMPDIR:-/tmp}/sol-rs-fixture
LVP_ICD=${LVP_ICD:-$(ls /usr/share/vulkan/icd.d/lvp_icd*.json | head -1)}
rm -rf "$WORK" && cp -r "$REF" "$WORK" && cd "$WORK"

capture() {  # example, width, height, frame to capture, extra args...
    local ex=$1 w=$2 h=$3 n=$4; shift 4
    sed -i -E "s/resolution: \[[0-9]+, [0-9]+\]/resolution: [$w, $h]/" "examples/$ex.rs"
    mkdir -p "$WORK/shots/$ex"
    ( cd "$WORK/shots/$ex" && VK_ICD_FILENAMES="$LVP_ICD" VK_INSTANCE_LAYERS=VK_LAYER_LUNARG_screenshot VK_SCREENSHOT_FRAMES=$n \
        timeout 600 xvfb-run -s "-screen 0 ${w}x${h}x24" cargo run --manifest-path "$WORK/Cargo.toml" --release --example "$ex" -- "$@" || true )
    ls "$WORK/shots/$ex"/*.ppm
}

# primary-hit ids (3-ray-debug with the id-writing stages)
cp "$HERE/debug_ids.rchit" assets/glsl/debug.rchit
cp "$HERE/debug_ids.rmiss" assets/glsl/debug.rmiss
capture 3-ray-debug 900 600 1
python "$HERE/ppm_to_fixture.py" ids "$WORK/shots/3-ray-debug/1.ppm" "$REPO/tests/golden/ref_ids_Duck_900x600.npz"

# accumulated path-traced frames (5-pathtrace, reference literals: 8 spp, 32 bounces)
capture 5-pathtrace 512 512 7 --model models/cornell.gltf
python "$HERE/ppm_to_fixture.py" frame "$WORK/shots/5-pathtrace/7.ppm" "$REPO/tests/golden/ref_frame_cornell_512x512_f8.npz" 8 0
capture 5-pathtrace 480 270 7 --model models/tunnel.gltf --sky
python "$HERE/ppm_to_fixture.py" frame "$WORK/shots/5-pathtrace/7.ppm" "$REPO/tests/golden/ref_frame_tunnel_480x270_f8.npz" 8 1
echo "fixtures written under $REPO/tests/golden: run python -m pytest tests/test_oracle.py -k reference_fixtures"
