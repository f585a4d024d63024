"""Stand-in for the un-vendored ``safety_gym`` package: re-exports the oracle's
restatement of Engine (oracle/sg_engine.py) under the import path the reference
uses (main/envs/zone_envs/ZoneEnvBase.py:5)."""
