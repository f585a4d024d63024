#!/bin/bash
# Regenerate EVERY fixture of tests/golden/ in a scratch copy of the repo by running the generators over the real
# reference code (/root/reference; build container only) and compare them byte for byte with the committed ones.
#   bash tests/golden/regen_check.sh        -> "37 same", no DIFF line, exit 0
set -u
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/../.." && pwd)"
TMP="$(mktemp -d)"
cp -r "$ROOT/oracle" "$ROOT/tests" "$ROOT/combinatorial_rl_tasks_b200" "$ROOT/include" "$TMP/"
rm -f "$TMP"/tests/golden/*.npz
cd "$TMP" || exit 2
for g in "gen_golden.py" "gen_golden_gae.py" "gen_golden_goals.py" "gen_golden_hard.py main" "gen_golden_hard.py zone-goals" \
         "gen_golden_model.py" "gen_golden_walls.py"; do
  python tests/golden/$g > "$TMP/$(echo "$g" | tr ' ' '_').log" 2>&1 || { echo "FAILED: $g (log in $TMP)"; exit 1; }
done
bad=0; same=0
for f in "$ROOT"/tests/golden/*.npz; do
  if cmp -s "$f" "$TMP/tests/golden/$(basename "$f")"; then same=$((same + 1)); else echo "DIFF $(basename "$f")"; bad=$((bad + 1)); fi
done
echo "$same same, $bad different (scratch copy: $TMP)"
exit $bad
