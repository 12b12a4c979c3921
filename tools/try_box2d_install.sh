#!/bin/bash
# GPU box (or any box): tries every offline route to a real pybox2d and logs the outcome.
# Output: gpurun_out/box2d_install.log (copied to profiles/box2d_install_rNN.log by hand).
out=${1:-gpurun_out/box2d_install.log}
mkdir -p "$(dirname "$out")"
{
  echo "== $(date -u +%FT%TZ) host $(hostname)"
  echo "== python -c 'import Box2D'";           python -c 'import Box2D; print(Box2D.__version__)' 2>&1 | tail -1
  echo "== python -c 'import gym'";             python -c 'import gym; print(gym.__version__)' 2>&1 | tail -1
  echo "== which swig";                         which swig || echo "swig: not found"
  echo "== wheelhouse";                         ls /opt/wheelhouse 2>/dev/null | grep -i -E 'box2d|gym|swig|pygame' || echo "no box2d/gym/swig/pygame wheel in /opt/wheelhouse"
  echo "== find / -iname '*box2d*'";            find / -xdev -iname '*box2d*' -not -path '*/repo/*' -not -path "$PWD/*" 2>/dev/null | head -5; echo "(end of find)"
  echo "== pip install --no-index --find-links /opt/wheelhouse --target baseline/_ref box2d-py"
  timeout 120 python -m pip install --no-index --find-links /opt/wheelhouse --target baseline/_ref box2d-py 2>&1 | tail -4
  echo "== pip download box2d-py (index; expected to fail: no network)"
  timeout 60 python -m pip download --no-deps -d /tmp/b2dl box2d-py 2>&1 | tail -3
  echo "== pip install box2d-py (index)"
  timeout 60 python -m pip install --target baseline/_ref box2d-py 2>&1 | tail -3
  echo "== result"; PYTHONPATH=baseline/_ref python -c 'import Box2D; print("Box2D importable", Box2D.__version__)' 2>&1 | tail -1
} > "$out" 2>&1
cat "$out"
