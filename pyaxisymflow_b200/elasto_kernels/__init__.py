"""Mirror of the reference package ``pyaxisymflow.elasto_kernels`` (same module and callable names)."""
