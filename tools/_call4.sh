bash tools/round_check.sh r01c
bash tools/collect_profiles.sh r01c
