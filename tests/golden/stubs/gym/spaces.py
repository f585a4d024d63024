from oracle.spaces import Box, Dict  # noqa: F401
