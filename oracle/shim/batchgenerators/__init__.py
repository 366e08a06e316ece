# ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement shim of an un-vendored dependency
# (nnunet@77bc485 / batchgenerators==0.21).  Never imported by the product path.
