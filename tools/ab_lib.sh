# A/B of two builds of librnb.so on one box, alternating processes: build/librnb_prev.so (an earlier commit, built by hand
# from `git archive`) against the in-tree library. Usage: bash tools/ab_lib.sh [arch batch]
mkdir -p gpurun_out
A=${1:-resnet50}; B=${2:-256}
for i in 1 2 3; do
  RNB_LIB=build/librnb_prev.so python tools/ab.py $A $B "" 2>&1 | grep "burst" | sed 's/^/prev: /'
  python tools/ab.py $A $B "" 2>&1 | grep "burst" | sed 's/^/new:  /'
done
