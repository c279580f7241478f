"""Reference arm of the benchmark: the UNMODIFIED reference env, run on the host cores.

``baseline/_ref/`` (git-ignored, shipped to the GPU box with the snapshot) receives a verbatim copy of the
reference's env modules when ``__graft_entry__.build()`` runs in a container that has ``/root/reference``;
``baseline/ref_env.py`` imports them from there behind stand-ins for three absent third-party packages.
Nothing under ``jolineedle_b200/`` imports this package.
"""
