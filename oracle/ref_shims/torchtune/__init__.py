"""TEST-ONLY shim package for `torchtune` (pinned 0.6.1, uv.lock:4511-4512); see modules.py."""
