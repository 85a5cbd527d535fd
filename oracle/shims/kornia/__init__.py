"""Import shim (TEST INFRASTRUCTURE ONLY): frido.modules.encoders.modules imports kornia at module scope."""
