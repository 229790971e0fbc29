"""Mirror of ``pyaxisymflow.utils`` for the pieces that sit on either side of the timestep loop
(SURVEY.md 8f-4): field output.  Plotting helpers and level-set shape generators are out of scope."""
