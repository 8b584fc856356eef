cd "$(dirname "$0")/.."
for n in 120000 200000 400000 1000000; do timeout 120 python tools/loop_debug.py $n 2>&1 | tail -1; done
echo "--- n=1M variants"
FPSB_LOOP=1 timeout 120 python tools/loop_debug.py 1000000 2>&1 | tail -1
FPSB_LOOP_NSPEC=0 timeout 120 python tools/loop_debug.py 1000000 2>&1 | tail -1
FPSB_LOOP_NSPEC=2 timeout 120 python tools/loop_debug.py 1000000 2>&1 | tail -1
FPSB_LOOP_CHUNK=2 timeout 120 python tools/loop_debug.py 1000000 2>&1 | tail -1
FPSB_LOOP_CHUNK=1 timeout 120 python tools/loop_debug.py 400000 2>&1 | tail -1
