for spec in "layer4.1.conv1 34 0" "layer4.1.conv1 148 0" "layer3.1.conv1 7 0" "layer2.1.conv1 2 0" "layer1.0.conv1 1 0" "layer4.1.conv1 34 2"; do
  set -- $spec
  python tools/profile_target.py 50 $1 $2 0 $3
done
