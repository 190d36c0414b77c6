"""Known-answer tests of the Fortran interpreter (oracle/ftn/) that runs the reference's sources for the pins.

The interpreter is test infrastructure, and the pins are only as good as its reading of Fortran: each test below is a small
program whose result follows from the language standard (and, for rounding, from IEEE-754 arithmetic) and is computed here
independently in Python / numpy."""
import math
import os
import struct

import numpy as np
import pytest

from oracle.ftn.interp import Interp


def run(src, tmp_path, files=None):
    p = tmp_path / "prog.f90"
    p.write_text(src)
    for name, data in (files or {}).items():
        (tmp_path / name).write_bytes(data if isinstance(data, bytes) else data.encode())
    I = Interp(cwd=str(tmp_path)).load([str(p)])
    stop = I.run_program()
    assert stop is None, stop
    return I


def out(I):
    return [l.strip() for l in I.io.stdout_lines]


def test_arithmetic_kinds_and_rounding(tmp_path):
    I = run("""
module m
  implicit none
  real(8):: a, b, c, d, e, p3, p5, pm2, s1, s2
  real(4):: r4
  integer:: i1, i2, i3, i4, i5
end module
program t
  use m
  implicit none
  real(8):: x(5)
  a = 0.1            ! default-real literal: single precision, then converted
  b = 0.1d0
  c = 1/3            ! integer division
  d = 1.d0/3
  e = 2.5d0 * 3 + 7/2
  r4 = 16777216.0 + 1.0     ! single precision addition rounds
  i1 = 7/2;  i2 = -7/2;  i3 = mod(-7,3);  i4 = nint(2.5d0);  i5 = int(-2.7d0)
  p3 = 1.1d0**3;  p5 = 1.1d0**5;  pm2 = 3.d0**(-2)
  x = [1.d16, 1.d0, -1.d16, 1.d0, 1.d0]
  s1 = sum(x)                      ! sequential: ((((1e16+1)-1e16)+1)+1) = 2
  s2 = x(1) + (x(2) + x(3)) + x(4) ! parentheses kept
end program
""", tmp_path)
    v = I.modules["m"].vars
    assert v["a"] == np.float64(np.float32(0.1)) and v["a"] != 0.1
    assert v["b"] == 0.1
    assert v["c"] == 0.0 and v["d"] == 1.0 / 3.0
    assert v["e"] == 2.5 * 3 + 3
    assert v["r4"] == np.float32(16777216.0)
    assert (v["i1"], v["i2"], v["i3"], v["i4"], v["i5"]) == (3, -3, -1, 3, -2)
    x = np.float64(1.1)
    assert v["p3"] == x * x * x
    assert v["p5"] == x * ((x * x) * (x * x))          # libgcc's __powidf2 for n = 5: x * (x^2)^2
    assert v["pm2"] == 1.0 / 9.0
    assert v["s1"] == 2.0
    assert v["s2"] == (1e16 + (1.0 - 1e16)) + 1.0
    assert isinstance(v["a"], np.float64) and isinstance(v["r4"], np.float32) and type(v["i1"]) is int


def test_arrays_bounds_sections_and_intrinsics(tmp_path):
    I = run("""
module m
  implicit none
  integer, parameter:: n = 4
  integer, parameter:: ix(3) = [3, 1, 2]
  real(8):: a(0:n,2), b(3), mm(2,2), mv(2), dp, tr(2,3)
  integer:: lo, hi, cnt, ml(1), after
  real(8), allocatable:: w(:,:)
  logical:: was
end module
program t
  use m
  implicit none
  integer:: i, j
  real(8):: r(2,3), v(3)
  do j = 1, 2
    do i = 0, n
      a(i,j) = 10*j + i
    enddo
  enddo
  after = i                       ! loop variable after completion: n+1
  lo = lbound(a,1); hi = ubound(a,1)
  b = a(1:3,2)
  b(ix) = b                       ! vector subscript on the left
  was = allocated(w)
  allocate(w(0:1,3))
  w = 1.5d0
  w(1,:) = [1.d0, 2.d0, 3.d0]
  r = reshape([1.d0,2.d0,3.d0,4.d0,5.d0,6.d0],[2,3])   ! column major
  v = [1.d0, 10.d0, 100.d0]
  mv = matmul(r, v)
  mm = matmul(r, transpose(r))
  dp = dot_product(v, r(2,:))
  tr = r
  cnt = count(r > 2.5d0)
  ml = maxloc(v)
  deallocate(w)
end program
""", tmp_path)
    v = I.modules["m"].vars
    assert v["after"] == 5 and (v["lo"], v["hi"]) == (0, 4)
    assert v["a"].lb == (0, 1) and v["a"].d[3, 1] == 23.0
    assert v["b"].d.tolist() == [22.0, 23.0, 21.0]          # b(3)=21, b(1)=22, b(2)=23
    assert v["was"] is False and v["w"] is None
    r = np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]])
    assert v["mv"].d.tolist() == [531.0, 642.0]
    assert np.array_equal(v["mm"].d, r @ r.T)
    assert v["dp"] == 642.0 and v["cnt"] == 4 and v["ml"].d.tolist() == [3]


def test_procedures_arguments_and_types(tmp_path):
    I = run("""
module shapes
  implicit none
  type :: box
    real(8):: w = 2.d0, h
    integer:: hits
    real(8), allocatable:: q(:)
  contains
    procedure:: area => box_area
    procedure:: grow => box_grow
  end type
  type(box):: boxes(2)
  real(8):: res(8)
  integer:: calls = 0
contains
  function box_area(this) result(a)
    class(box), intent(in):: this
    real(8):: a
    a = this%w * this%h
  end function
  subroutine box_grow(this, f, extra)
    class(box), intent(inout):: this
    real(8), intent(in):: f
    real(8), intent(in), optional:: extra
    this%w = this%w * f
    this%hits = this%hits + 1
    if (present(extra)) this%h = this%h + extra
    calls = calls + 1
  end subroutine
  subroutine swap(a, b)
    real(8):: a, b, t
    t = a; a = b; b = t
  end subroutine
  subroutine fill(f, nz, ny, val)
    integer, intent(in):: nz, ny
    real(8), intent(inout):: f(0:nz-1, ny)      ! explicit shape: new bounds on the caller's storage
    real(8), intent(in):: val
    f(0, ny) = val
  end subroutine
  recursive function fact(n) result(r)
    integer, intent(in):: n
    integer:: r
    if (n <= 1) then
      r = 1
    else
      r = n * fact(n-1)
    endif
  end function
  subroutine host(x, y)
    real(8):: x, y, k
    k = 3.d0
    call inner(x)
    y = twice(x)
  contains
    subroutine inner(z)
      real(8):: z
      z = z * k           ! host association
    end subroutine
    function twice(z)
      real(8):: z, twice
      twice = 2.d0 * z
    end function
  end subroutine
end module
program t
  use shapes
  implicit none
  real(8):: p, q, g(3,2)
  type(box):: c
  integer:: counter
  boxes(1)%h = 5.d0
  call boxes(1)%grow(1.5d0)
  call boxes(1)%grow(2.d0, extra=1.d0)
  res(1) = boxes(1)%area()
  c = boxes(1)               ! value copy
  c%w = 100.d0
  res(2) = boxes(1)%w
  p = 1.d0; q = 2.d0
  call swap(p, q)
  res(3) = p; res(4) = q
  g = 0.d0
  call fill(g, 3, 2, 7.d0)
  res(5) = g(1,2)
  res(6) = fact(5)
  call host(p, q)
  res(7) = p; res(8) = q
  boxes(:)%hits = 9
end program
""", tmp_path)
    v = I.modules["shapes"].vars
    assert v["res"].d.tolist() == [6.0 * 6.0, 6.0, 2.0, 1.0, 7.0, 120.0, 6.0, 12.0]
    assert v["calls"] == 2
    assert [b.f["hits"] for b in v["boxes"].d] == [9, 9]


def test_characters_and_formatted_output(tmp_path):
    I = run("""
program t
  implicit none
  character(len=10):: s, f
  character(len=3):: b
  character(len=40):: line
  integer:: i, n
  real(8):: x
  s = 'Ab'
  s = adjustr(s)
  write(*,'(A)') '['//s//']'
  write(f,'(I10)') 42
  do i = 1, 10
    if (f(i:i) == ' ') f(i:i) = '0'
  enddo
  write(*,'(A)') f
  write(b,'(I3)') 7
  b = adjustr(b)
  write(*,'(A,I8,A,F14.8)') ' Steps:', 12, '  Time/Tref:', 0.125d0
  write(*,'(A,F18.12)') ' FIELDSTAT L2 u ', 1.001329469031d0
  write(*,'(3(I2,1X),E12.4)') 1, 2, 3, 12345.678d0
  write(*,'(ES12.4)') 0.000123456d0
  line = '  3 4.5d0, hello  '
  read(line,*) n, x, s
  write(*,'(I2,F6.2,A)') n, x, trim(s)//'|'
  write(*,'(I0,A,L1)') index('abcdef','cd'), ' ', ('abc' == 'abc   ')
  write(*,'(A)') achar(iachar('a') - 32)//trim('xy  ')//'.'
  write(*,'(I3)') len_trim('  ab  ')
end program
""", tmp_path)
    o = I.io.stdout_lines
    assert o[0] == "[        Ab]"
    assert o[1] == "0000000042"
    assert o[2] == " Steps:      12  Time/Tref:    0.12500000"
    assert o[3] == " FIELDSTAT L2 u     1.001329469031"
    assert o[4] == " 1  2  3   0.1235E+05"
    assert o[5] == "  1.2346E-04"
    assert o[6] == " 3  4.50hello|"
    assert o[7] == "3 T"
    assert o[8] == "Axy."
    assert o[9] == "  4"


def test_files_list_directed_and_stream(tmp_path):
    data = struct.pack("<ii", 2, 7) + struct.pack("<d", 0.5) + np.arange(6, dtype="<f8").tobytes()
    I = run("""
module m
  implicit none
  integer:: nb, st, ios, k, cnt
  real(8):: tm, a(2,3), tot
  character(len=20):: word
  logical:: there, nothere
end module
program t
  use m
  implicit none
  character(len=256):: buffer
  integer:: i
  inquire(file='in.bin', exist=there)
  inquire(file='nope.bin', exist=nothere)
  open(unit=13, file='in.bin', form='unformatted', status='old', access='stream')
  read(13) nb, st, tm
  read(13) a
  close(13)
  open(unit=14, file='out.bin', form='unformatted', access='stream')
  write(14) nb+1, tm*2
  write(14) a(2,:)
  close(14)
  open(unit=111, file='in.txt', status='old', action='read')
  cnt = 0
  tot = 0.d0
  ios = 0
  do while (ios == 0)
    read(111, '(a)', iostat=ios) buffer
    if (ios /= 0) exit
    buffer = adjustl(buffer)
    if (buffer(1:1) == '#') cycle
    read(buffer, *) k, word
    cnt = cnt + 1
    tot = tot + k
  enddo
  close(111)
  open(unit=15, file='out.txt')
  write(15,'(A,I4)') 'count', cnt
  write(15,'(A)', advance='no') 'a='
  write(15,'(F6.1)') tot
  close(15)
end program
""", tmp_path, files={"in.bin": data, "in.txt": "# header\n 1 one\n  # skipped\n 20 twenty\n300 three  extra\n"})
    v = I.modules["m"].vars
    assert (v["nb"], v["st"], v["tm"]) == (2, 7, 0.5)
    assert v["a"].d.tolist() == [[0.0, 2.0, 4.0], [1.0, 3.0, 5.0]]       # column-major fill
    assert v["there"] is True and v["nothere"] is False
    assert (v["cnt"], v["tot"], v["word"].strip()) == (3, 321.0, "three")
    raw = (tmp_path / "out.bin").read_bytes()
    assert raw == struct.pack("<i", 3) + struct.pack("<d", 1.0) + np.array([1.0, 3.0, 5.0]).tobytes()
    assert (tmp_path / "out.txt").read_text() == "count   3\na= 321.0\n"


def test_control_flow(tmp_path):
    I = run("""
module m
  implicit none
  integer:: r(6)
end module
program t
  use m
  implicit none
  integer:: i, j, s
  s = 0
  do i = 10, 1, -3          ! 10 7 4 1
    s = s + i
  enddo
  r(1) = s; r(2) = i        ! i = -2 after the loop
  s = 0
  do i = 1, 5
    do j = 1, 5
      if (j > i) exit
      if (mod(j,2) == 0) cycle
      s = s + 1
    enddo
  enddo
  r(3) = s
  select case (r(1))
  case (1:10)
    r(4) = 1
  case (22)
    r(4) = 2
  case default
    r(4) = 3
  end select
  i = 0
  do while (.true.)
    i = i + 1
    if (i*i > 50) exit
  enddo
  r(5) = i
  if (i < 3) then
    r(6) = 1
  elseif (i < 9) then
    r(6) = 2
  else
    r(6) = 3
  endif
end program
""", tmp_path)
    assert I.modules["m"].vars["r"].d.tolist() == [22, -2, 9, 2, 8, 2]


def test_math_intrinsics_match_libm(tmp_path):
    I = run("""
module m
  implicit none
  real(8):: v(10)
end module
program t
  use m
  implicit none
  real(8):: x
  x = 0.7d0
  v(1) = dsin(x); v(2) = cos(x); v(3) = dsqrt(x); v(4) = x**1.5d0; v(5) = dacos(x)
  v(6) = dabs(-x); v(7) = max(x, 0.2d0, 0.9d0); v(8) = sign(x, -1.d0); v(9) = floor(-x); v(10) = exp(x)
end program
""", tmp_path)
    x = 0.7
    want = [math.sin(x), math.cos(x), math.sqrt(x), math.pow(x, 1.5), math.acos(x), x, 0.9, -x, -1.0, math.exp(x)]
    assert I.modules["m"].vars["v"].d.tolist() == want


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference sources are not on this machine")
def test_reference_sources_parse_and_load():
    """Every source file of the reference's Makefile:33 parses; the module variables the hot path reads come out as declared."""
    from oracle.ftn.interp import load_reference
    I = load_reference()
    m = I.modules["constparams"]
    I.ready(m)
    assert m.vars["lbmdim"] == 18
    ee = m.vars["ee"]
    assert ee.lb == (0, 1) and ee.d.shape == (19, 3) and ee.d[7].tolist() == [1, 1, 0]
    assert m.vars["wt"].d[0] == 1.0 / 3.0 and m.vars["oppo"].d.tolist()[7] == 10


def test_integer_powers_match_libgcc_powidf2():
    """real**integer beyond x**2 is a call of libgcc's __powidf2 in a gfortran build without fast-math: the interpreter's chain of
    multiplications must give the same bits as the libgcc of this machine."""
    import ctypes
    from oracle.ftn import rt
    try:
        fn = ctypes.CDLL("libgcc_s.so.1").__powidf2
    except (OSError, AttributeError):
        pytest.skip("libgcc_s.so.1 / __powidf2 not available")
    fn.restype, fn.argtypes = ctypes.c_double, [ctypes.c_double, ctypes.c_int]
    rng = np.random.default_rng(7)
    for x in np.concatenate([rng.uniform(0.01, 3.0, 200), -rng.uniform(0.01, 3.0, 50)]):
        for n in range(1, 13):
            assert float(rt.fpow(np.float64(x), n)) == fn(float(x), n), (x, n)
        for n in (-1, -2, -3, -5):
            assert float(rt.fpow(np.float64(x), n)) == fn(float(x), n), (x, n)
