package RMGpu;
# Loader for the XS glue (RMGpu.xs) over librmgpu.so.  Built by `make` in this directory.
use strict;
use warnings;
require XSLoader;
our $VERSION = '0.1';
XSLoader::load( 'RMGpu', $VERSION );
1;
