timeout 200 python profiles/phase_profile.py 2>&1 | grep "us/launch"
