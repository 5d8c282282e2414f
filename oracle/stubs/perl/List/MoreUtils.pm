# Test-infrastructure shim: the reference's minimath generator wants
# List::MoreUtils::pairwise, which is not installed in this image. This provides
# just that one function so the UNMODIFIED generator script can be run where it
# lies under /root/reference (see oracle/Makefile).
package List::MoreUtils;
use strict;
use warnings;
use Exporter 'import';
our @EXPORT_OK = qw(pairwise);

sub pairwise(&\@\@)
{
    my ($f, $x, $y) = @_;
    my $caller = caller;
    my @r;
    no strict 'refs';
    for my $i (0 .. ($#$x > $#$y ? $#$x : $#$y))
    {
        local (*{"${caller}::a"}, *{"${caller}::b"}) = (\$x->[$i], \$y->[$i]);
        push @r, $f->();
    }
    return @r;
}
1;
