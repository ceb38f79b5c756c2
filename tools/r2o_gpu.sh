python tools/sc_round_profile.py 20 2>&1 | tail -24
