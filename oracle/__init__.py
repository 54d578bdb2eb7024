"""oracle — TEST INFRASTRUCTURE ONLY. See oracle/README.md. Never imported by renderer_b200."""
