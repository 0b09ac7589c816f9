cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "acquire_track or fam5_tracking or b1c_wb or b1c_nb" 2>&1 | grep -v "^\[parity\]\|window" | tail -25
