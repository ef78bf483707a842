#!/bin/sh
# UTAU / OpenUtau resampler launcher of goofer_b200 (same 13 arguments as SillySampler.sh -> SillySampler.py).
# Copy this file and goofer_b200.yaml (python -m goofer_b200.manifest > goofer_b200.yaml) into the Resamplers folder.
# With no arguments it starts the HTTP front-end on port 8572, like the reference.
HERE=$(cd "$(dirname "$0")"; pwd -P)
PYTHONPATH="$HERE/..${PYTHONPATH:+:$PYTHONPATH}" exec "${GOOFER_PYTHON:-python3}" -m goofer_b200.cli "$@"
