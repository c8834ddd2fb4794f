"""Empty stand-in: utils/data.py imports tensorflow_datasets at module level (utils/__init__.py:1 pulls it in);
nothing of it is on the tested path.  TEST INFRASTRUCTURE ONLY."""
