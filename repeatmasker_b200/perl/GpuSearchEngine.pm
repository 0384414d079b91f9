#!/usr/bin/perl
##---------------------------------------------------------------------------##
##  File: GpuSearchEngine.pm
##
##  A SearchEngineI plug-in for RepeatMasker: the GPU sibling of
##  CrossmatchSearchEngine.pm.  Instead of popen()ing the external cross_match
##  binary and parsing its text (CrossmatchSearchEngine.pm:444, :549-879) it
##  calls librmgpu.so (B200, sm_100a) through the XS glue in RMGpu.xs and builds
##  the same SearchResultCollection from flat hit records.
##
##  It ISA CrossmatchSearchEngine so that every `isa("CrossmatchSearchEngine")`
##  branch of the driver keeps taking the crossmatch path (matrix directory
##  RepeatMasker:1987, FASTA libraries :7631/:8193, cache glob :7879) and the
##  class method parseOutput() -- the universal .cat/.out/.align reader used by
##  RepeatMasker:1429/4009/4762 and ProcessRepeats:9284-9392 -- is inherited
##  untouched.
##
##  Install: see INTEGRATION.md (two edits in RepeatMasker, PERL5LIB entry).
##---------------------------------------------------------------------------##

=head1 NAME

GpuSearchEngine - RepeatMasker search engine backed by librmgpu.so (B200)

=head1 SYNOPSIS

  use GpuSearchEngine;
  my $engine = GpuSearchEngine->new( pathToEngine => "/path/to/librmgpu.so" );
  $engine->setMatrix( ".../Matrices/crossmatch/20p43g.matrix" );
  $engine->setQuery( "batch.masked" );
  $engine->setSubject( "lib.fa" );
  my ( $status, $resultCollection ) = $engine->search( minMatch => 9, bandwidth => 14 );

=cut

package GpuSearchEngine;
use strict;
use Carp;
use File::Basename;
use SearchEngineI;
use SearchResultCollection;
use SearchResult;
use CrossmatchSearchEngine;
use RMGpu;

use vars qw(@ISA $VERSION);
@ISA     = qw(CrossmatchSearchEngine);
$VERSION = 0.1;

my $CLASS = "GpuSearchEngine";

##---------------------------------------------------------------------------##
## Constructor.  CrossmatchSearchEngine::new croaks without pathToEngine and
## executes it (CrossmatchSearchEngine.pm:188-260); here pathToEngine is
## optional and names the shared library (only used for the .cat/.tbl trailer,
## RepeatMasker:1267-1270).  No CUDA call is made here: the engine object is
## built in the parent and copied into every forked batch child
## (RepeatMasker:602, :1043); the CUDA context is created lazily inside
## search() in the child.
##---------------------------------------------------------------------------##
sub new {
  my $class          = shift;
  my %nameValuePairs = @_;

  my $this = {};
  bless $this, $class;
  $this->{'device'} = -1;
  if ( defined $ENV{'RMGPU_DEVICE'} && $ENV{'RMGPU_DEVICE'} =~ /^\d+$/ ) {
    $this->{'device'} = $ENV{'RMGPU_DEVICE'};
  }

  my $path = delete $nameValuePairs{'pathToEngine'};
  $path = _defaultLibraryPath() if ( !defined $path || $path eq "" );
  $this->setPathToEngine( $path );
  _ensureServer( $path );

  if ( %nameValuePairs ) {
    while ( my ( $name, $value ) = each( %nameValuePairs ) ) {
      my $method = "set" . ucfirst( $name );
      unless ( $this->can( $method ) ) {
        croak( "$CLASS" . "::set: Instance variable $name doesn't exist." . "" );
      }
      $this->$method( $value );
    }
  }
  return $this;
}

sub _defaultLibraryPath {
  my $dir = dirname( $INC{'GpuSearchEngine.pm'} || __FILE__ );
  return "$dir/../librmgpu.so";
}

##---------------------------------------------------------------------------##
## The persistent GPU worker.  RepeatMasker builds ONE engine object in the
## parent (RepeatMasker:602) and forks a short-lived child per batch
## (RepeatMasker:1040-1180); a CUDA context cannot cross fork, and a context
## plus a packed library per child would cost more than the searches.  So the
## constructor -- still without any CUDA call in this process -- starts
## rmgpu_server (all visible GPUs, library resident, pending calls of the
## children coalesced into one pass per GPU) unless RMGPU_SERVER already names
## one, and exports its socket to the children through the environment.  The
## process that started it stops it when it exits.  RMGPU_NO_SERVER=1 keeps the
## in-process path (one CUDA context per search() call).
##---------------------------------------------------------------------------##
my ( $serverPid, $serverOwner, $serverSocket );

sub _ensureServer {
  my $libPath = shift;
  return if ( defined $ENV{'RMGPU_SERVER'} && $ENV{'RMGPU_SERVER'} ne "" );
  return if ( $ENV{'RMGPU_NO_SERVER'} );
  my $bin = dirname( $libPath ) . "/rmgpu_server";
  return unless ( -x $bin );
  my $sock = ( $ENV{'TMPDIR'} || "/tmp" ) . "/rmgpu.$$.sock";
  my $gpus = defined $ENV{'RMGPU_GPUS'} ? $ENV{'RMGPU_GPUS'} : "all";
  my $pid  = fork();
  return unless ( defined $pid );
  if ( $pid == 0 ) {
    open( STDERR, ">", "/dev/null" ) unless ( $ENV{'RMGPU_SERVER_LOG'} );
    exec( $bin, "--socket", $sock, "--gpus", $gpus, "--die-with-parent" ) or exit( 1 );
  }
  for ( my $i = 0 ; $i < 400 ; $i++ ) {
    last if ( -S $sock && RMGpu::server_ping( $sock ) == 0 );
    if ( waitpid( $pid, 1 ) == $pid ) {    # WNOHANG: the worker died (no GPU): keep the in-process path, which reports it
      return;
    }
    select( undef, undef, undef, 0.025 );
  }
  $ENV{'RMGPU_SERVER'} = $sock;
  ( $serverPid, $serverOwner, $serverSocket ) = ( $pid, $$, $sock );
}

END {
  if ( defined $serverPid && $$ == $serverOwner ) {
    my $status = $?;
    RMGpu::server_quit( $serverSocket );
    waitpid( $serverPid, 0 );
    $? = $status;
  }
}

sub setPathToEngine {
  my $this  = shift;
  my $value = shift;
  my $old   = $this->{'pathToEngine'};
  $this->{'pathToEngine'} = $value;
  my $banner = RMGpu::version();
  if ( $banner =~ /rmgpu version ([\.\d]+)/ ) {
    $this->{'version'} = $1;
  }
  return $old;
}

sub setDevice {
  my $this = shift;
  my $old  = $this->{'device'};
  $this->{'device'} = shift;
  return $old;
}

sub getDevice {
  my $this = shift;
  return $this->{'device'};
}

##---------------------------------------------------------------------------##
## getParameters() is inherited (CrossmatchSearchEngine.pm:277-363): pathToEngine
## names an existing file (librmgpu.so), so its engine check passes and the
## driver's error messages (RepeatMasker:2076, :2101) read
## "<path>/librmgpu.so -alignments -gap_init ...".
##---------------------------------------------------------------------------##

##---------------------------------------------------------------------------##
## search(): same contract as CrossmatchSearchEngine.pm:381-524 --
##   ( $status, $SearchResultCollection ) = $engine->search( minMatch=>, bandwidth=> )
## status 0 is success; a non-zero status walks the driver's retry ladder
## (RepeatMasker:2068-2107).
##---------------------------------------------------------------------------##
sub search {
  my $this           = shift;
  my %nameValuePairs = @_;

  if ( %nameValuePairs ) {
    while ( my ( $name, $value ) = each( %nameValuePairs ) ) {
      my $method = "set" . ucfirst( $name );
      unless ( $this->can( $method ) ) {
        croak( $CLASS . "::search: Instance variable $name doesn't exist." );
      }
      $this->$method( $value );
    }
  }

  my $cmdLine = $this->getParameters();    # inherited: croaks on missing inputs

  my $matrixName;
  {
    my @path = split( /[\\\/]/, $this->getMatrix() );
    $matrixName = $path[ $#path ];
    $matrixName =~ s/[\n\r]//g;
  }

  print $CLASS . "::search(): Running librmgpu as:\n    $cmdLine\n"
      if ( $this->getDEBUG() );

  my $raw = 0;
  $raw = 1
      if (    $this->getScoreMode()
           && $this->getScoreMode() == SearchEngineI::basicScoreMode );

  # cross_match defaults for anything the driver did not set
  my $gapInit   = defined $this->getGapInit()   ? $this->getGapInit()   : -4;
  my $insGapExt = defined $this->getInsGapExt() ? $this->getInsGapExt() : -3;
  my $delGapExt = defined $this->getDelGapExt() ? $this->getDelGapExt() : -3;
  my $minScore  = $this->getMinScore()  || 30;
  my $maskLevel = defined $this->getMaskLevel() ? $this->getMaskLevel() : 80;
  my $minMatch  = $this->getMinMatch()  || 14;
  my $bandwidth = $this->getBandwidth() || 14;

  my ( $status, $error, $records ) = RMGpu::search(
    $this->getQuery(),  $this->getSubject(), $this->getMatrix(), $gapInit,
    $insGapExt,         $delGapExt,          $minScore,          $maskLevel,
    $minMatch,          $bandwidth,          $raw,               ( $this->getWordRaw() ? 1 : 0 ),
    ( $this->getGenerateAlignments() ? 1 : 0 ), $this->{'device'}
  );
  if ( $status ) {
    warn $CLASS . "::search(): $error\n";
    return ( $status, SearchResultCollection->new() );
  }

  my $searchResultsCollection =
      recordsToCollection( $records, $matrixName );
  return ( 0, $searchResultsCollection );
}

##---------------------------------------------------------------------------##
## recordsToCollection(): flat hit records -> SearchResultCollection, with the
## exact post-steps of the crossmatch path:
##   - field set of the hit line (CrossmatchSearchEngine.pm:599-606, :746-765;
##     a trailing '*' lands in 'id' exactly as parseOutput leaves it);
##   - strand C: the rows arrive as cross_match displays them (query reverse
##     complemented) and are flipped so the query is forward (:824-837);
##   - pctDelete re-based on the subject span from the 2-decimal text (:497-503);
##   - Kimura divergence with the CpG modification (:509-511).
##---------------------------------------------------------------------------##
sub recordsToCollection {
  my $records    = shift;
  my $matrixName = shift;

  my $resultColl = SearchResultCollection->new();
  foreach my $rec ( @{$records} ) {
    my %nameValuePairs = (
                           score          => $rec->{'score'},
                           pctDiverge     => $rec->{'pctDiverge'},
                           pctDelete      => $rec->{'pctDelete'},
                           pctInsert      => $rec->{'pctInsert'},
                           queryName      => $rec->{'queryName'},
                           queryStart     => $rec->{'queryStart'},
                           queryEnd       => $rec->{'queryEnd'},
                           queryRemaining => $rec->{'queryRemaining'},
                           orientation    => $rec->{'orientation'},
                           subjName       => $rec->{'subjName'},
                           subjStart      => $rec->{'subjStart'},
                           subjEnd        => $rec->{'subjEnd'},
                           subjRemaining  => $rec->{'subjRemaining'},
                           id             => ( $rec->{'star'} ? "*" : undef ),
                           overlap        => undef
    );
    my $result = GpuSearchResult->new( %nameValuePairs );

    if ( defined $rec->{'queryString'} ) {
      my $querySeq = $rec->{'queryString'};
      my $subjSeq  = $rec->{'subjString'};
      if ( $rec->{'orientation'} eq "C" ) {
        $querySeq = reverse $querySeq;
        $querySeq =~ tr/ACGTYRMKHBVD/TGCARYKMDVBH/;
        $subjSeq = reverse $subjSeq;
        $subjSeq =~ tr/ACGTYRMKHBVD/TGCARYKMDVBH/;
      }
      $result->setQueryString( $querySeq );
      $result->setSubjString( $subjSeq );
    }
    $result->setMatrixName( $matrixName ) if ( defined $matrixName && $matrixName ne "" );
    $resultColl->add( $result );

    my $qryLen = $result->getQueryEnd() - $result->getQueryStart() + 1;
    my $sbjLen = $result->getSubjEnd() - $result->getSubjStart() + 1;
    if ( $qryLen != 0 && $sbjLen != 0 ) {
      my $deletions = ( $result->getPctDelete() / 100 ) * $qryLen;
      $result->setPctDelete( sprintf( "%4.2f", ( $deletions * 100 ) / $sbjLen ) );
    }
    if ( defined $rec->{'pctKimuraDiverge'} ) {
      # computed by the library with the arithmetic of SearchResult::calcKimuraDivergence( divCpGMod => 1 ):
      # at GPU search rates the per-character Perl loop would cost more than the search itself
      $result->setPctKimuraDiverge( $rec->{'pctKimuraDiverge'} );
    }
    elsif ( defined $rec->{'queryString'} ) {
      my ( $div, $transi, $transv, $wellCharBases, $numCpGs ) =
          $result->calcKimuraDivergence( divCpGMod => 1 );
      $result->setPctKimuraDiverge( sprintf( "%4.2f", $div ) );
    }
  }
  return $resultColl;
}

##---------------------------------------------------------------------------##
## GpuSearchResult: the SearchResult the engine hands out.  Same object, same
## fields; only the query-oriented alignment text -- what the batch children
## write into the .cat file, RepeatMasker:1489 -> SearchResultCollection::write
## -> toStringFormatted( AlignWithQuerySeq ) -- comes from the library
## (rmgpu_format_alignment) instead of the per-column Perl loop of
## SearchResult.pm:2639-2834.  The annotation line is the inherited routine's.
## It is rendered when asked for, from the fields the object holds at that
## moment, so everything the driver changes in between (names, offsets of a
## fragment, ids) shows.  Every other format, rows of unequal length and rows
## with a character outside the nucleotide codes take the inherited routine.
##---------------------------------------------------------------------------##
package GpuSearchResult;
use strict;
use SearchResult;
use vars qw(@ISA);
@ISA = qw(SearchResult);

sub _toCrossMatchFormat {
  my $obj           = shift;
  my $alignmentMode = shift;
  my $displayParams = shift;

  my $query   = $obj->getQueryString();
  my $subject = $obj->getSubjString();
  if (
       defined $alignmentMode
    && (    $alignmentMode == SearchResult::AlignWithQuerySeq
         || $alignmentMode == SearchResult::N_AlignWithQuerySeq )
    && defined $query
    && defined $subject
    && $query ne ""
    && length( $query ) == length( $subject )
    && $query !~ /[^ACGTBDHVRYKMSWNXacgtbdhvrykmswnx-]/
    && $subject !~ /[^ACGTBDHVRYKMSWNXacgtbdhvrykmswnx-]/
      )
  {
    # the annotation line alone, with the id order of the requested mode
    my $header = $obj->SUPER::_toCrossMatchFormat(
                          $alignmentMode == SearchResult::AlignWithQuerySeq
                          ? SearchResult::NoAlign
                          : SearchResult::N_NoAlign
    );
    my $block = RMGpu::format_alignment(
                          $obj->getQueryName(),  $obj->getSubjName(),
                          $query,                $subject,
                          $obj->getQueryStart(), $obj->getQueryEnd(),
                          $obj->getSubjStart(),  $obj->getSubjEnd(),
                          ( $obj->getOrientation() eq "C" ? 1 : 0 ),
                          $obj->getMatrixName(), $obj->getPctKimuraDiverge(),
                          $obj->getCpGSites(),   $obj->getPctRawKimuraDiverge()
    );
    return $header . $block if ( defined $block );
  }
  return $obj->SUPER::_toCrossMatchFormat( $alignmentMode, $displayParams );
}

1;
