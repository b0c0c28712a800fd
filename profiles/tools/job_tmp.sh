timeout 300 python profiles/tools/fused_phase_profile.py
