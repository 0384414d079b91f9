#!/usr/bin/perl
# CPU-side check of the Perl glue: the collection that the reference's own CrossmatchSearchEngine->search() builds
# from the oracle's cross_match text must equal the collection GpuSearchEngine::recordsToCollection() builds from
# flat hit records of the same search (records passed as JSON by the caller, in the RMGpu::search() layout).
# usage: compare_collections.pl <cross_match binary> <matrix> <query> <subject> <records.json>
use strict;
use JSON::PP;
use CrossmatchSearchEngine;
use GpuSearchEngine;
use SearchResult;

my ( $bin, $matrix, $query, $subject, $json ) = @ARGV;
my $ref = CrossmatchSearchEngine->new( pathToEngine => $bin );
$ref->setMatrix( $matrix ); $ref->setQuery( $query ); $ref->setSubject( $subject );
$ref->setGenerateAlignments( 1 ); $ref->setGapInit( -30 ); $ref->setInsGapExt( -6 ); $ref->setDelGapExt( -5 );
$ref->setMaskLevel( 101 ); $ref->setMinScore( 225 );
$ref->setScoreMode( SearchEngineI::complexityAdjustedScoreMode );
my ( $status, $collA ) = $ref->search( minMatch => 9, bandwidth => 14 );
die "reference engine status $status" if $status;

local $/;
open my $fh, "<", $json or die $!;
my $records = decode_json( <$fh> );
my @p = split( /\//, $matrix );
my $collB = GpuSearchEngine::recordsToCollection( $records, $p[-1] );

die "size " . $collA->size() . " vs " . $collB->size() . "\n" if $collA->size() != $collB->size();
my $n = 0;
for ( my $i = 0 ; $i < $collA->size() ; $i++ ) {
  my ( $a, $b ) = ( $collA->get( $i ), $collB->get( $i ) );
  my $sa = $a->toStringFormatted( SearchResult::AlignWithQuerySeq );
  my $sb = $b->toStringFormatted( SearchResult::AlignWithQuerySeq );
  die "hit $i differs:\n$sa\n---\n$sb\n" if $sa ne $sb;
  foreach my $m ( qw(getPctKimuraDiverge getPctDelete getPctInsert getPctDiverge getOrientation getMatrixName getId) ) {
    my ( $x, $y ) = ( $a->$m(), $b->$m() );
    $x = "<undef>" unless defined $x; $y = "<undef>" unless defined $y;
    die "hit $i $m: $x vs $y\n" if $x ne $y;
  }
  $n++;
}
print "OK $n\n";
