"""WaveGlow vocoder surface of the reference (WaveGlow/Modules.py names), running on libmstts_b200."""
