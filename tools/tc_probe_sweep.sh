cd $GRAFT_REPO_ROOT
P="timeout 120 python tests/tc_probe.py"
echo "== wgrad only (K-major/K-major)"; $P --mask 3 2>&1 | tail -12
echo "== fwd only default"; $P --mask 6 2>&1 | tail -12
echo "== dgrad only default"; $P --mask 5 2>&1 | tail -12
echo "== fwd V1 swap lbo/sbo"; $P --mask 6 --small --knobs 2=512,3=4096 2>&1 | tail -3
echo "== fwd V2 sw128 plain"; $P --mask 6 --small --knobs 1=2,5=3,2=4096,3=1024 2>&1 | tail -3
echo "== fwd V3 sw128 plain swapped"; $P --mask 6 --small --knobs 1=2,5=3,2=1024,3=4096 2>&1 | tail -3
