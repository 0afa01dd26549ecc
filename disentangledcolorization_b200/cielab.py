"""CIELAB ab-gamut quantisation table (313 bins of 10x10 in the ab plane).

Replaces `utils/cielab.py:5-64` (`ABGamut`, `CIELAB.q_to_ab`) of the reference, which loads
`utils/gamut_pts.npy` through a cwd-relative path.  The 313 in-gamut bins of Zhang et al. (ECCV'16)
form 20 contiguous b-intervals, one per a-value, so the table is stored here as run-lengths and
expanded at import time; `q_to_ab[q]` is the bin centre, ordered by (a, b) exactly as
`ab[ab_gamut_mask] + AB_BINSIZE/2` orders it (`utils/cielab.py:62-64`).
"""
import numpy as np

AB_BINSIZE = 10
N_BINS = 313

# (a, b_lo, b_hi) -- inclusive, step 10
_GAMUT_ROWS = (
    (-90, 50, 90), (-80, 20, 90), (-70, 0, 90), (-60, -20, 90), (-50, -30, 100),
    (-40, -40, 100), (-30, -50, 100), (-20, -50, 100), (-10, -60, 100), (0, -70, 100),
    (10, -80, 90), (20, -80, 90), (30, -90, 90), (40, -100, 90), (50, -100, 80),
    (60, -110, 80), (70, -110, 80), (80, -110, 70), (90, -110, 70), (100, -90, 0),
)


def q_to_ab():
    """(313, 2) float32 bin centres in Lab units (not divided by 110)."""
    pts = [(a, b) for a, lo, hi in _GAMUT_ROWS for b in range(lo, hi + AB_BINSIZE, AB_BINSIZE)]
    out = np.asarray(pts, dtype=np.float32)
    assert out.shape == (N_BINS, 2)
    return out


Q_TO_AB = q_to_ab()


# Empirical prior of the 313 bins (ImageNet ab statistics of Zhang et al.), the reference's `utils/gamut_probs.npy`
# (313 float64, loaded by `ABGamut`, utils/cielab.py:5-17), stored as base85 of the little-endian float64 bytes.  Used by
# the training-side class re-balancing only (`ColorLabel.weights`, models/basic.py:155-157).
_PRIOR_B85 = (
    "O-=yd#ye_0YMr}NaF>EUVDR;hzF4C^pnmLz?<2-O5E}O&ALFk+X^Pc`R3>9SSjzIjjRuiEDC8o{F+iz4`A8Y2+;qS`*X9!yKT63y#2E?!$bQs5"
    "4b9ardZ^bv$c9Df@wT--Ip1rM&?I?2CQpfx_)(=k=_6u2BiO<|8T*2+77Nxsr}bbCKw{!P4(!jW{T}Q-UkPMPaVqmZ6N_JqZQAfY^x%%EmP+0}"
    "7kfAFCH=ELZdWw&ZycOH!<as1Q>?!}69Omi!&KKkF81M%qdo0DM`Qalwm<nk_$$r!#Wn&z01#{14@n6>-n~VSVDS+@C0_orV)7C{<5K6$RE7pW"
    ")~w}rFG=w};Fb-Kn_<a5k(MWp@Aj!a2wCiFxIXbdX?(lnt^fx=oW=47C(8~$<Xol%1l|@uk3ZPzk0KvG2we|9ljbEqY_1F1q(Cu0ffD1cWiK{A"
    "SvEiV<qt7GyR+$2lcXR&!OMW-{cr_8ByT3lF;U(=eJQBnZOw{4;M<!c3h%8xc0|A5RMYl8k*+CspMD)bR**;&eq1a+z;m0RKyNWWj_o=ZtHw1y"
    "Cb6tBu9HANoEZ7Im48k@w~TMCgN0T<UYd^{(hXEU^%z{NwQ)v2bO8P*PINIpR7dbJ0(%rc*uC{d6}ay{HG0GYbT_9yD=V59Pg1Hr-tKNBHc$IL"
    "U#Ungf3zk)z5J63EABf#8@CEvq`*c$%|p2khtf?yP)l%-;aFBb<?K)rRKjIHM%Ki)s*Q0!iIZ(hP~vbuZ={GJE-GX{kFPTyKT%OXBk2W`VemIU"
    "X#eo9cb6VNq6MUk*`EYIynnI-Od{7l0(6}0piKBa2pX1%-mfe_g{$9X9`Z&%XZVY0OYc@cxn6ZxA~$3|iI#(4*5z<NI^_Bv1xI^7YKh+^18ajn"
    "oxJ1M&1Hi>gHE-<VWN0H|MgNRerjevxmNH`7+X<4HWi+eF4sFhq&(@0r4%baj;&;V-F_WE6K*jI!5;HIo}~Ki>Nx8@zhfF@GXp9=@i1Q9&cskZ"
    "^B4Tf@4sk22NYF+y#9MXcObMuQk|1O(gGfP%_ONmZ@`+?vmT>AAEs^1*$$RJar8?*J$;Bj{%n-w4Fq;S`>%F(pD<!Sm>h}e;cHMonbv5V<%2;#"
    "y>suXO_no1uMRZVMzkG2LnIhnCHLJv^o{5F>0;GBtp+(V;H(}$9pvohe(*^@<iY;lK3r!%&a#%DD=B+Fh*FD!UQCZafj<ySns2Z_J;}~I3Bt)g"
    "pon67bGf@eQk`o$eL$r@w$P@&-tv$?9z<Qo`bBv^tX3Crvu0#Je_(8KE;Lg=fSb6o4~0ZOw-kf+L-I5~8VWBpkOC7w%6u|Vj8eKjl0+|(()G4J"
    "gkHCa_BIJWw#xX)u|PLJ8rq0)N260eNy$Rq+~8wBOOC%yB=l}Skp11F+7x>~IU5Q!lLwSPS&qsx2QjNZxHx$RlftM!DE~@TCc>aUY~(9t{SuNt"
    "kJeOS$p(KvG(ogzCNyb33FFaLfss`|4Q%N#{qsXVGO(aUf^;xH(q^4_%vJwB5CH5$J{R#mW$JkE!@naxNM(yRsc1bvY1KlZRh3FVov7SUdfQMx"
    "BOrwvtcy`Ua}d$`7MD{$bo4n<5A<O_<YU;4x-E7;c}V#qF^Gjfe`DlMun~zru1t=lh9!hQ+|fYZkdt>mp+Yj5xYuSsBZ4ok7$a0ai8@s6yZ=Bx"
    "pOv1;ivA=&|G7RFicsP{*XXgQF4@jLZn4f34~`E%!2TusmX9evnIoTODv&il3yOy-f3i70h=Wd{`wTWeFRYH-&Db(OwqGE1Y*IHrm&s7|U#UYs"
    "iwh1|O<7exB7-$wE*oe+iwMbdr`B;l*pIMhAy#oe$ir#oZgFWpct3W>>r!4nvDjnuF=b9aD~)v%i2OM}3V!t*uT~d7p4|)#T{FTyItyi$Pb;H7"
    "&2>r@3R?9(OJ3WpVVD>{{ltagL>nYOhB2fpqP8YKJ;AIlWK|+RB3VrQE}|PhGqXcaj6xtk7$APd#2qF-h2hBxDc><a6I*CqNjE@0@wk#X5Jpfx"
    "H=!uEZaH2*$cV0pe*I!UX<6H5WuRR@)V}p)i~LeQ2AIO~FQG&~S|=6|uZ1x`1+<LKSKkCbie6K9s8O6gkCvXZ{^!&_Kj_(Av>*pR^MjhBe6bZj"
    "Z}G+&X3!KrY_(*o6MhUo&qjBL$)5#3+wi$}@PP?GPjSf%u7eRj=RumzxiT0(*enyR?1m#hWhDHrI|ML4J?=bp-&;ICFPE!xRnAI3zTn!D6wX#Z"
    "-e9-Y8d+IC6d+0LM*mJfop4Mnr%pXT2C+;U>Ln&W?6Gu7a53;cOHVMna;~gCaTW-$#hdXy)hQI&xt0b$rvhnc5OoAUB^#5P9pCjnH6!K}@&oWb"
    "=?<KAl#B8{5ld}9uR!}gcqqQc))50gvi{b=<5vtns$}Tst40<-a^6oKtAZdu=z$agIC?2RQSbV|oI*E0I4cj!$e=|(B-Bx)$;ePYFBVf)a^z1x"
    "c035ECzLxsFlCb|TeloPtWbE?O`6v}vQX!h%7fTG`o{8CXVv*Wqxryl#=!DEc%7Dwq6FYRu@){eLm<{Z9#gQ_=$_a<u7pA(ozUYxaj$e6)2i)0"
    "jPK|)TjldU<4<&8KehcndXB@^W#|PzR8-1|2m}p3eyeZxV>S~%9A-CMqsbmWVHT1NZh|a7T#@#eoq0GvOAq+DtR+D|iFhlFu--X8883e5Tj3W!"
    "7Aoumz|X@zp1l0kK{4Mx(yK&3*oovmtfrL)2;a><sfQOFPQb!Gn8VMqf3Ctl&igsLW<bb3=#ms`6~xv)+H#c4$(P$cPGM6;0r%iO%BL*FIi2P{"
    "dYlS3Tm|kvz;u;!l=kpGv)ig{7kTnNO7!2zj8gYL_a0K1LHq$fou!dIu)YmHR@uvulQ|ea7BB<}9+DwH{?VA8$p;TVeQt+k_G+0v>Fj-LjliBh"
    "837(f!Y80Ug`7pfaXzO$6=(ocHXN}&ZS)dO^0LD|W$TdOX^O}`<(+R2`NYRQj-1*xZsNo~`)3)<WOl(mc(dE&vd+Oi>iQnMTI{<%xHBVo#-O!6"
    "MeLGjDzUOY<dJLKub;F&P8E8>@5;J9&TQfhQ?0~4@fsd9f=1OocpY<_IJ3?^?#@ibpA$MhfMkVWHlSlZr!NRP-guKfK5j!pagVn?6yQ1KS|+4E"
    "H6oa&_QsJu&M~){Dms8Zyx_j#(y(bh*&ZDJukc<zU;{ggp(Rf~"
)


def gamut_prior():
    import base64
    out = np.frombuffer(base64.b85decode(_PRIOR_B85), dtype="<f8").copy()
    assert out.shape == (N_BINS,)
    return out


def class_weights(lambda_=0.5):
    """`ColorLabel.weights` (models/basic.py:153-157): w = 1 / ((1 - lambda) prior + lambda uniform), normalised so that
    sum(prior * w) == 1.  float32: the reference's ABGamut casts the table to float32 (utils/cielab.py:8-11)."""
    prior = gamut_prior().astype(np.float32)
    uniform = np.zeros_like(prior)
    uniform[prior > 0] = np.float32(1.0) / np.float32((prior > 0).sum())
    w = np.float32(1.0) / (np.float32(1.0 - lambda_) * prior + np.float32(lambda_) * uniform)
    return (w / np.sum(prior * w, dtype=np.float32)).astype(np.float32)
