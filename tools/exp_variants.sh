#!/usr/bin/env bash
# builds scan-kernel variants with experiment switches and times the forward scan (diagnostics only)
set -e
for v in "-DCB_EXP_NOSTORE" "-DCB_EXP_NOLOAD2" "-DCB_EXP_NOSTORE -DCB_EXP_NOLOAD2 -DCB_EXP_NOSTEP"; do
  CB200_EXTRA_NVCC="$v" python -m consenrich_b200.build --force > /dev/null
  echo "== variant [$v]"
  python - <<PY
import sys; sys.argv=["x"]
exec(open("tools/phase_probe.py").read().replace("(4, 8, 11, 16)","(16,)"))
PY
done
python -m consenrich_b200.build --force > /dev/null
