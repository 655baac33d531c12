for spec in "layer4.1.conv1 148 0" "layer4.1.conv1 148 2" "layer4.1.conv1 74 0" "layer4.1.conv1 96 0"; do
  set -- $spec
  timeout 120 python tools/profile_target.py 300 $1 $2 1 $3
done 2>&1
for bn in 16 32 48 64; do echo "ADMMQ_FORCE_BN=$bn"; ADMMQ_FORCE_BN=$bn timeout 120 python tools/profile_target.py 300 layer4.1.conv1 148 1 0; done 2>&1
