for cfg in "100 220 2000" "100 128 2000" "150 220 1000" "100 220 400"; do
  python tools/chain_probe.py $cfg 2>&1 | tail -1
  DISTGCN_B200_LIB=$PWD/distgcn_b200/libdistgcn_b200_half.so python tools/chain_probe.py $cfg 2>&1 | tail -1
done
