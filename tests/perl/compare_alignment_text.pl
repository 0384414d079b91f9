#!/usr/bin/perl
# CPU-side check of GpuSearchResult (GpuSearchEngine.pm): the alignment text rendered by librmgpu
# (rmgpu_format_alignment) must equal, byte for byte, what the reference's SearchResult::toStringFormatted prints for
# an object with the same fields -- in both query-oriented modes, for every other mode (inherited), and over the edge
# cases of the format: chunks that are all gap, ambiguity codes, lower case, long names, optional fields present or
# undef, orientation C, rows of 1 / 50 / 51 columns, coordinates wider than the field.
# usage: compare_alignment_text.pl [<alignments.json>]   (a list of {q, s, strand} objects; random rows are always added)
use strict;
use JSON::PP;
use Time::HiRes qw(time);
use SearchResult;
use GpuSearchEngine;

my @cases;
if ( @ARGV ) {
  local $/;
  open my $fh, "<", $ARGV[ 0 ] or die $!;
  my $list = decode_json( <$fh> );
  foreach my $a ( @{$list} ) {
    push @cases, [ $a->{'q'}, $a->{'s'}, ( $a->{'strand'} ? "C" : "" ) ];
  }
}

# hand-made edge cases
push @cases, [ "A", "A", "" ], [ "A", "G", "C" ], [ "-", "A", "" ], [ "A", "-", "C" ];
push @cases, [ "ACGT" x 12 . "AC", "ACGT" x 12 . "AT", "" ];             # exactly 50 columns
push @cases, [ "ACGT" x 12 . "ACG", "ACGT" x 12 . "ATG", "C" ];          # 51 columns
push @cases, [ "ACGTACGTAC" . "-" x 120 . "ACGTNRYKM", "ACGTACGTAC" . "ACGTTGCAAC" x 12 . "ACGTACGTA", "" ];   # all-gap chunks (query)
push @cases, [ "ACGTACGTAC" . "ACGTTGCAAC" x 12 . "ACGTNRYKM", "ACGTACGTAC" . "-" x 120 . "ACGTACGTA", "C" ];  # all-gap chunks (subject)
push @cases, [ "acgtnACGTNxXrRyY-bdhv", "ACGTNacgtnXxYyRr-BDHA", "" ], [ "RYKMSWBDHVNX", "AAAAAAAAAAAA", "C" ];
push @cases, [ "AC--GT", "ACTTGT", "" ], [ "--ACGT", "TTACGT", "" ], [ "ACGT--", "ACGTTT", "C" ];

# random rows over the full alphabet, gaps never in both rows
srand( 20261020 );
my @alpha = ( ( "A", "C", "G", "T" ) x 12, qw(R Y K M S W B D H V N X a c g t n x) );
for my $t ( 1 .. 300 ) {
  my $len = 1 + int( rand( $t < 250 ? 400 : 4000 ) );
  my ( $q, $s ) = ( "", "" );
  my $gap = 0;
  for ( 1 .. $len ) {
    my ( $a, $b ) = ( $alpha[ rand @alpha ], $alpha[ rand @alpha ] );
    $b = $a if ( rand() < 0.7 );
    my $r = rand();
    if ( $gap > 0 ) { $gap--; $r = $gap & 1 ? 0.0 : 0.03; }
    elsif ( $r < 0.06 ) { $gap = int( rand( 70 ) ); }
    if    ( $r < 0.03 ) { $a = "-"; }
    elsif ( $r < 0.06 ) { $b = "-"; }
    $q .= $a;
    $s .= $b;
  }
  push @cases, [ $q, $s, ( $t & 1 ? "C" : "" ) ];
}

my ( $n, $tRef, $tNew ) = ( 0, 0, 0 );
my $i = 0;
foreach my $c ( @cases ) {
  my ( $q, $s, $orient ) = @{$c};
  $i++;
  my $qBases = ( $q =~ tr/-//c );
  my $sBases = ( $s =~ tr/-//c );
  my $qStart = ( $i % 7 == 0 ) ? 12345678901 : 1 + ( $i * 37 ) % 60000;
  my $sStart = 1 + ( $i * 11 ) % 3000;
  my %f = (
            score          => 200 + $i,
            pctDiverge     => "12.34",
            pctDelete      => "1.20",
            pctInsert      => "0.50",
            queryName      => ( $i % 3 == 0 ? "chr1_frag-$i and a description" : $i % 3 == 1 ? "q$i" : "averyveryverylongqueryname$i" ),
            queryStart     => $qStart,
            queryEnd       => $qStart + ( $qBases > 0 ? $qBases - 1 : 0 ),
            queryRemaining => 1000 + $i,
            orientation    => $orient,
            subjName       => ( $i % 2 ? "AluY#SINE/Alu" : "L1PA3_3end#LINE/L1" ),
            subjStart      => $sStart,
            subjEnd        => $sStart + ( $sBases > 0 ? $sBases - 1 : 0 ),
            subjRemaining  => 5 + $i,
            id             => ( $i % 4 == 0 ? "*" : $i % 4 == 1 ? $i : undef ),
            overlap        => ( $i % 5 == 0 ? "*" : undef ),
            lineageId      => ( $i % 6 == 0 ? "m_b1s001i$i" : undef ),
  );
  my ( $a, $b ) = ( SearchResult->new( %f ), GpuSearchResult->new( %f ) );
  foreach my $o ( $a, $b ) {
    $o->setQueryString( $q );
    $o->setSubjString( $s );
    $o->setMatrixName( "25p41g.matrix" )        if ( $i % 3 );
    $o->setPctKimuraDiverge( "13.57" )           if ( $i % 4 );
    $o->setCpGSites( 7 )                         if ( $i % 5 == 1 || $i % 5 == 2 );
    $o->setPctRawKimuraDiverge( "15.01" )        if ( $i % 5 == 2 || $i % 5 == 3 );
  }
  foreach my $mode ( SearchResult::AlignWithQuerySeq, SearchResult::N_AlignWithQuerySeq ) {
    my $t0 = time();
    my $sa = $a->toStringFormatted( $mode );
    my $t1 = time();
    my $sb = $b->toStringFormatted( $mode );
    my $t2 = time();
    $tRef += $t1 - $t0;
    $tNew += $t2 - $t1;
    die "case $i mode $mode differs:\n$sa\n---\n$sb\n" if ( $sa ne $sb );
    $n++;
  }
  foreach my $mode ( SearchResult::NoAlign, SearchResult::AlignWithSubjSeq, SearchResult::OutFileFormat ) {
    my ( $sa, $sb ) = ( $a->toStringFormatted( $mode ), $b->toStringFormatted( $mode ) );
    die "case $i mode $mode (inherited) differs\n" if ( $sa ne $sb );
  }
  # the object is rendered from what it holds NOW: a change the driver makes after the search must show
  $a->setQueryName( "renamed" ); $b->setQueryName( "renamed" );
  $a->setQueryStart( $a->getQueryStart() + 60000 ); $b->setQueryStart( $b->getQueryStart() + 60000 );
  $a->setQueryEnd( $a->getQueryEnd() + 60000 ); $b->setQueryEnd( $b->getQueryEnd() + 60000 );
  die "case $i after setters differs\n"
      if ( $a->toStringFormatted( SearchResult::AlignWithQuerySeq ) ne $b->toStringFormatted( SearchResult::AlignWithQuerySeq ) );
  # a clone keeps the class (the driver clones results)
  my $cl = $b->clone();
  die "case $i: clone is a " . ref( $cl ) . "\n" unless ( ref( $cl ) eq "GpuSearchResult" );
}
# rows the library does not take (unequal length, a character outside the nucleotide codes) fall back to the inherited routine
{
  my %f = ( score => 1, pctDiverge => "0.00", pctDelete => "0.00", pctInsert => "0.00", queryName => "q", queryStart => 1,
            queryEnd => 4, queryRemaining => 0, orientation => "", subjName => "s", subjStart => 1, subjEnd => 4, subjRemaining => 0 );
  foreach my $rows ( [ "ACGT", "ACG" ], [ "ACGU", "ACGT" ], [ "AC.T", "ACGT" ] ) {
    my ( $a, $b ) = ( SearchResult->new( %f ), GpuSearchResult->new( %f ) );
    local $SIG{__WARN__} = sub { };
    foreach my $o ( $a, $b ) { $o->setQueryString( $rows->[ 0 ] ); $o->setSubjString( $rows->[ 1 ] ); }
    die "fallback case differs\n"
        if ( $a->toStringFormatted( SearchResult::AlignWithQuerySeq ) ne $b->toStringFormatted( SearchResult::AlignWithQuerySeq ) );
  }
}
printf "OK %d texts, reference %.3f s, library %.3f s\n", $n, $tRef, $tNew;
