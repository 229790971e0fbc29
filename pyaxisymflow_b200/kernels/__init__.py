"""Mirror of the reference package ``pyaxisymflow.kernels`` (same module and callable names)."""
