for g in 33 148; do timeout 120 python tools/profile_target.py 300 layer4.1.conv1 $g 1 2; done 2>&1
for g in 33 36 148; do timeout 120 python tools/profile_target.py 300 layer4.1.conv1 $g 1 0; done 2>&1
timeout 120 python tools/profile_target.py 300 layer3.1.conv1 7 1 0
