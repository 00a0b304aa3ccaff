"""Mirror of the reference's `segmentation/ops` package (modules/ + functions/): put this package's
parent on sys.path under the name `ops` and `from ops.modules import
MSDeformAttn` resolves to the B200 implementation."""
