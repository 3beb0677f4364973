"""Reference-side bindings: what a MIND maintainer adds to run the unmodified simulator / planner on this library
(INTEGRATION.md).  Nothing here is on the measured path."""
