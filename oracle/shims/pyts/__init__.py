"""Stand-in for pyts==0.12.0 (environment.yml:118 of the reference); TEST INFRASTRUCTURE ONLY."""
