"""CPU models of the training kernels (kfnet_b200/csrc/train.cu) checked against torch autograd of the oracle's forward
functions -- no GPU.  The CUDA kernels are line-by-line transcriptions of these models:
  * build_ops / run_gemm / wgrad: the operand re-packing descriptors (tap offsets, strides, phases) and row-grid geometry of
    the forward, data-gradient and weight-gradient GEMMs of every layer kind (conv s1 / s2 / 1x1, conv2d_transpose, dense);
  * one unrolled filter step backwards (filter_backward_kernel) and the loss forward + gradient (train_loss_kernel);
  * soft-argmax / sigma_w, difference volume and l2-normalise backward."""
import sys

import numpy as np
import torch

from oracle import kfnet_oracle as O
from oracle import kfnet_train_oracle as TO

D = torch.float64
rng = np.random.default_rng(1)

def build_ops(kind,k,cin,cout,stride):
    """python port of train.cu build_ops: returns (fwd ops, dgrad ops); op = dict(A,B,sa,sb,tapoff,dy,dx,istride,omul,oaddy,oaddx)"""
    def base(A,B,sa,sb): return dict(A=A,B=B,sa=sa,sb=sb,tapoff=[],dy=[],dx=[],istride=1,omul=1,oaddy=0,oaddx=0)
    fwd=[];dg=[]
    if kind=="dense" or (kind=="conv" and k==1):
        f=base(cin,cout,cout,1); f["tapoff"]=[0]; f["dy"]=[0]; f["dx"]=[0]; fwd.append(f)
        d=base(cout,cin,1,cout); d["tapoff"]=[0]; d["dy"]=[0]; d["dx"]=[0]; dg.append(d); return fwd,dg
    if kind=="conv":
        pad=1 if stride==1 else 0
        f=base(cin,cout,cout,1); f["istride"]=stride
        for ky in range(3):
            for kx in range(3):
                t=ky*3+kx; f["tapoff"].append(t*cin*cout); f["dy"].append(ky-pad); f["dx"].append(kx-pad)
        fwd.append(f)
        if stride==1:
            d=base(cout,cin,1,cout)
            for t in range(9): d["tapoff"].append(t*cin*cout); d["dy"].append(-(t//3-1)); d["dx"].append(-(t%3-1))
            dg.append(d)
        else:
            for ph in range(4):
                py,px=ph>>1,ph&1
                d=base(cout,cin,1,cout); d["omul"]=2; d["oaddy"]=py; d["oaddx"]=px
                kys=[1 if py else 0,2]; dys=[0,-1]; kxs=[1 if px else 0,2]; dxs=[0,-1]
                for a in range(1 if py else 2):
                    for c in range(1 if px else 2):
                        d["tapoff"].append((kys[a]*3+kxs[c])*cin*cout); d["dy"].append(dys[a]); d["dx"].append(dxs[c])
                dg.append(d)
        return fwd,dg
    for ph in range(4):
        py,px=ph>>1,ph&1
        f=base(cin,cout,1,cin); f["omul"]=2; f["oaddy"]=py; f["oaddx"]=px
        kys=[1 if py else 0,2]; dys=[0,-1]; kxs=[1 if px else 0,2]; dxs=[0,-1]
        for a in range(1 if py else 2):
            for c in range(1 if px else 2):
                f["tapoff"].append((kys[a]*3+kxs[c])*cin*cout); f["dy"].append(dys[a]); f["dx"].append(dxs[c])
        fwd.append(f)
    d=base(cout,cin,cin,1); d["istride"]=2
    for t in range(9): d["tapoff"].append(t*cin*cout); d["dy"].append(t//3); d["dx"].append(t%3)
    dg.append(d)
    return fwd,dg

def run_gemm(op,W,inp,out,MH,MW,accum=False):
    """conv_simt semantic: inp [N,IH,IW,A] -> out [N,OH,OW,B] (in place) over the row grid"""
    N,IH,IW,_=inp.shape
    Wf=W.reshape(-1)
    for t in range(len(op["tapoff"])):
        M=np.zeros((op["A"],op["B"]))
        for a in range(op["A"]):
            for b in range(op["B"]): M[a,b]=Wf[op["tapoff"][t]+a*op["sa"]+b*op["sb"]]
        for my in range(MH):
            for mx in range(MW):
                iy=my*op["istride"]+op["dy"][t]; ix=mx*op["istride"]+op["dx"][t]
                if 0<=iy<IH and 0<=ix<IW:
                    out[:,my*op["omul"]+op["oaddy"],mx*op["omul"]+op["oaddx"],:]+=inp[:,iy,ix,:]@M

def wgrad(inp,dout,MH,MW,istride,dys,dxs,A,B):
    N,IH,IW,_=inp.shape
    dW=np.zeros((len(dys),A,B))
    for t in range(len(dys)):
        for my in range(MH):
            for mx in range(MW):
                iy=my*istride+dys[t]; ix=mx*istride+dxs[t]
                if 0<=iy<IH and 0<=ix<IW: dW[t]+=inp[:,iy,ix,:].T@dout[:,my,mx,:]
    return dW

def check(kind,k,cin,cout,stride,H,Wd):
    N=2
    if kind=="conv": wshape=(k,k,cin,cout)
    elif kind=="deconv": wshape=(3,3,cout,cin)
    else: wshape=(cin,cout)
    W=rng.normal(size=wshape); bvec=rng.normal(size=cout)
    x=torch.tensor(rng.normal(size=(N,H,Wd,cin)),dtype=D,requires_grad=True)
    Wt=torch.tensor(W,dtype=D,requires_grad=True); bt=torch.tensor(bvec,dtype=D)
    if kind=="conv": y=O.conv2d_same(x,Wt,bt,stride,False)
    elif kind=="deconv": y=O.conv2d_transpose_same(x,Wt,bt,False)
    else: y=O.conv2d_same(x,Wt.reshape(1,1,cin,cout),bt,1,False)
    gy=torch.tensor(rng.normal(size=tuple(y.shape)),dtype=D)
    (y*gy).sum().backward()
    fwd,dg=build_ops(kind,k,cin,cout,stride)
    OH,OW=y.shape[1:3]
    # forward
    out=np.zeros(tuple(y.shape))+bvec
    xn=x.detach().numpy(); gyn=gy.numpy()
    if kind=="deconv":
        for op in fwd: run_gemm(op,W,xn,out,H,Wd)
    else: run_gemm(fwd[0],W,xn,out,OH,OW)
    ef=np.abs(out-y.detach().numpy()).max()
    # dgrad
    dx=np.zeros(xn.shape)
    if kind=="conv" and stride==2:
        for op in dg: run_gemm(op,W,gyn,dx,OH,OW)
    else: run_gemm(dg[0],W,gyn,dx,H,Wd)
    ed=np.abs(dx-x.grad.numpy()).max()
    # wgrad
    if kind=="deconv":
        d=dg[0]; dW=wgrad(gyn,xn,H,Wd,2,d["dy"],d["dx"],cout,cin).reshape(wshape)
    else:
        f=fwd[0]; dW=wgrad(xn,gyn,OH,OW,f["istride"],f["dy"],f["dx"],cin,cout).reshape(wshape)
    ew=np.abs(dW-Wt.grad.numpy()).max()
    print(kind,k,cin,cout,stride,(H,Wd),"fwd %.1e dgrad %.1e wgrad %.1e"%(ef,ed,ew))
    assert max(ef,ed,ew)<1e-9



def test_gemm_op_descriptors_forward_dgrad_wgrad():
    check("conv", 3, 3, 5, 1, 6, 8)
    check("conv", 3, 4, 6, 2, 6, 8)
    check("conv", 3, 4, 6, 2, 2, 2)
    check("conv", 1, 5, 3, 1, 4, 4)
    check("deconv", 3, 4, 3, 2, 3, 4)
    check("deconv", 3, 4, 3, 2, 1, 1)
    check("dense", 1, 6, 2, 1, 1, 1)


def test_filter_step_and_loss_backward_models():
    torch.manual_seed(0); rng = np.random.default_rng(0)
    GH,GW=8,12; P=GH*GW
    # ---------- filter step backward (numpy port of filter_backward_kernel) ----------
    prev=torch.randn(1,GH,GW,4,dtype=D); prev[...,3]=prev[...,3].abs()*0.1+0.05
    flow=(torch.rand(1,GH,GW,2,dtype=D)*7-4)
    sigw=torch.rand(1,GH,GW,1,dtype=D)*0.05+0.005
    z=torch.randn(1,GH,GW,4,dtype=D); z[...,3]=z[...,3].abs()*0.1+0.05
    Gt=torch.randn(1,GH,GW,4,dtype=D); Gk=torch.randn(1,GH,GW,4,dtype=D)
    pc=prev[...,:3].clone().requires_grad_(True); pu=prev[...,3:4].clone().requires_grad_(True)
    fl=flow.clone().requires_grad_(True); sw=sigw.clone().requires_grad_(True)
    zc=z[...,:3].clone().requires_grad_(True); zu=z[...,3:4].clone().requires_grad_(True)
    pm=O.get_pixel_map(1,GH,GW,D)+fl
    tc,tu=O.process_predict(pm, sw, pc, pu)
    kc,ku=O.build_kf_coord(tc,tu,zc,zu)
    L=(tc*Gt[...,:3]).sum()+(tu*Gt[...,3:4]).sum()+(kc*Gk[...,:3]).sum()+(ku*Gk[...,3:4]).sum()
    L.backward()
    # numpy port
    def port():
        G_prev=np.zeros((P,4)); dflow=np.zeros((P,2)); dsw=np.zeros(P); Gm=np.zeros((P,4))
        pv=prev.numpy().reshape(P,4); f=flow.numpy().reshape(P,2); s=sigw.numpy().reshape(P); zz=z.numpy().reshape(P,4)
        gt=Gt.numpy().reshape(P,4); gk=Gk.numpy().reshape(P,4)
        minvar=1e-10; xmax=GW-1; ymax=GH-1
        for p in range(P):
            y,x=divmod(p,GW); cx=x+f[p,0]; cy=y+f[p,1]
            x0=np.floor(cx); y0=np.floor(cy)
            x0s=min(max(x0,0),xmax); x1s=min(max(x0+1,0),xmax); y0s=min(max(y0,0),ymax); y1s=min(max(y0+1,0),ymax)
            wx0=x1s-cx; wx1=cx-x0s; wy0=y1s-cy; wy1=cy-y0s
            w=[wx0*wy0,wx0*wy1,wx1*wy0,wx1*wy1]
            idx=[int(y0s)*GW+int(x0s),int(y1s)*GW+int(x0s),int(y0s)*GW+int(x1s),int(y1s)*GW+int(x1s)]
            tap=[pv[i] for i in idx]
            xm=sum(w[i]*tap[i][:3] for i in range(4)); sl=sum(w[i]*tap[i][3] for i in range(4))
            sl2=sl*sl; sw2=s[p]*s[p]; lastv=max(sl2,minvar); tv=max(sw2,minvar); tu_=np.sqrt(tv+lastv)
            lv=tu_*tu_; mv=zz[p,3]**2; den=lv+mv; K=lv/den; omr=1-K; omk=max(omr,0); ku_=np.sqrt(omk*lv)
            d_xm=omk*gk[p,:3]+gt[p,:3]; d_z=K*gk[p,:3]
            d_omk=(gk[p,:3]*xm).sum(); d_K=(gk[p,:3]*zz[p,:3]).sum(); d_lv=0
            if ku_>0:
                h=gk[p,3]/(2*ku_); d_omk+=h*lv; d_lv+=h*omk
            if omr>=0: d_K-=d_omk
            d_lv+=d_K*mv/den**2; d_mv=-d_K*lv/den**2; d_sz=2*zz[p,3]*d_mv
            d_tu=gt[p,3]+2*tu_*d_lv; h2=d_tu/(2*tu_)
            d_sw=2*s[p]*h2 if sw2>=minvar else 0; d_sl=2*sl*h2 if sl2>=minvar else 0
            dw=[]
            for i in range(4):
                dw.append((d_xm*tap[i][:3]).sum()+d_sl*tap[i][3])
                G_prev[idx[i],:3]+=w[i]*d_xm; G_prev[idx[i],3]+=w[i]*d_sl
            d_wx0=dw[0]*wy0+dw[1]*wy1; d_wx1=dw[2]*wy0+dw[3]*wy1; d_wy0=dw[0]*wx0+dw[2]*wx1; d_wy1=dw[1]*wx0+dw[3]*wx1
            dflow[p]=[d_wx1-d_wx0,d_wy1-d_wy0]; dsw[p]=d_sw; Gm[p,:3]=d_z; Gm[p,3]=d_sz
        return G_prev,dflow,dsw,Gm
    G_prev,dflow,dsw,Gm=port()
    def e(a,b,n):
        err=float(np.abs(a-b).max()); print(n, err, float(np.abs(b).max())); assert err < 1e-9, (n, err)
    e(G_prev[:,:3],pc.grad.numpy().reshape(P,3),"d prev coord"); e(G_prev[:,3],pu.grad.numpy().reshape(P),"d prev unc")
    e(dflow,fl.grad.numpy().reshape(P,2),"d flow"); e(dsw,sw.grad.numpy().reshape(P),"d sigw")
    e(Gm[:,:3],zc.grad.numpy().reshape(P,3),"d z"); e(Gm[:,3],zu.grad.numpy().reshape(P),"d sigma_z")

    # ---------- loss kernel port ----------
    B=2
    est=torch.randn(B,GH,GW,4,dtype=D)*0.3; est[...,3]=est[...,3].abs()*0.2+0.02
    gt=est[...,:3]+torch.randn(B,GH,GW,3,dtype=D)*torch.where(torch.rand(B,GH,GW,1)>0.4,0.004,0.3)
    mask=(torch.rand(B,GH,GW,1)>0.25).to(D)
    img=torch.randint(0,256,(B,GH,GW,3)).to(D)   # already "resized"
    c=est[...,:3].clone().requires_grad_(True); u=est[...,3:4].clone().requires_grad_(True)
    loss,acc=TO.coord_loss_with_uncertainty(c,u,gt,mask)
    tot=loss+50*TO.smooth_loss(c,img,mask)
    tot.backward()
    def port_loss():
        E=est.numpy(); g=gt.numpy(); m=mask.numpy()[...,0]; I=img.numpy()
        valid=m.sum()+1; G=np.zeros_like(E); Ls=0; Ss=0; nz=0
        def iw(b,y0,x0,y1,x1): return np.exp(-0.625*np.abs(I[b,y0,x0]-I[b,y1,x1]).mean())
        for b in range(B):
          for y in range(GH):
            for x in range(GW):
                e_=E[b,y,x]; mm=m[b,y,x]; dv=e_[:3]-g[b,y,x]; d=(dv**2).sum(); uu=max(e_[3],1e-5)
                raw=3*np.log(uu)+d/(2*uu*uu); Ls+=mm*min(raw,-2.0); nz+= 1 if mm*d-0.0025>0 else 0
                gg=np.zeros(4)
                if raw<=-2.0:
                    gg[:3]=mm*dv/uu**2; gg[3]=mm*(3/uu-d/uu**3) if e_[3]>=1e-5 else 0
                h=np.zeros(3)
                if x+1<GW:
                    w=iw(b,y,x,y,x+1)*mm; a=e_[:3]-E[b,y,x+1,:3]; Ss+=(a**2).sum()/3*w; h+=a*w
                if x>0:
                    w=iw(b,y,x-1,y,x)*m[b,y,x-1]; h-=(E[b,y,x-1,:3]-e_[:3])*w
                if y+1<GH:
                    w=iw(b,y,x,y+1,x)*mm; a=e_[:3]-E[b,y+1,x,:3]; Ss+=(a**2).sum()/3*w; h+=a*w
                if y>0:
                    w=iw(b,y-1,x,y,x)*m[b,y-1,x]; h-=(E[b,y-1,x,:3]-e_[:3])*w
                gg[:3]=(gg[:3]+50*2/3*h)/valid; gg[3]/=valid
                G[b,y,x]=gg
        return G, Ls/valid+50*Ss/valid, (valid-nz)/valid
    G,Lv,Av=port_loss()
    e(G[...,:3],c.grad.numpy(),"loss d coord"); e(G[...,3:4],u.grad.numpy(),"loss d unc")
    assert abs(Lv-float(tot.detach()))<1e-12 and abs(Av-float(acc))<1e-12


def test_softmax_flow_diff_volume_l2norm_backward_models():
    # softmax-flow backward
    R=5
    logits=torch.tensor(rng.normal(size=(R,64))*3,dtype=D,requires_grad=True); unc=torch.tensor(rng.normal(size=(R,1)),dtype=D,requires_grad=True)
    offs=torch.tensor([[j-4,i-4] for i in range(8) for j in range(8)],dtype=D)
    prob=torch.softmax(logits,-1); flow=prob@offs; sig=torch.exp(unc)*1e-2
    gf=torch.tensor(rng.normal(size=(R,2)),dtype=D); gs=torch.tensor(rng.normal(size=(R,1)),dtype=D)
    ((flow*gf).sum()+(sig*gs).sum()).backward()
    p=prob.detach().numpy(); dl=np.zeros((R,64))
    for r in range(R):
        dp=np.array([gf[r,0]*((k&7)-4)+gf[r,1]*((k>>3)-4) for k in range(64)]); dot=(p[r]*dp).sum(); dl[r]=p[r]*(dp-dot)
    assert np.abs(dl-logits.grad.numpy()).max()<1e-12 and np.abs((gs*sig).detach().numpy()-unc.grad.numpy()).max()<1e-12
    # diff volume backward
    B,GH,GW=3,5,6; P=GH*GW
    f=torch.tensor(rng.normal(size=(B,GH,GW,32)),dtype=D,requires_grad=True)
    diffs=[]
    for j in range(B-1):
        d,_=O.coord_volume(f[j:j+1],f[j+1:j+2],8); diffs.append(d.reshape(-1,8,8,32))
    diff=torch.cat(diffs)   # [(B-1)P,8,8,32]
    g=torch.tensor(rng.normal(size=tuple(diff.shape)),dtype=D)
    (diff*g).sum().backward()
    gd=g.numpy().reshape(B-1,P,64,32); df=np.zeros((B,P,32))
    for fi in range(B):
        for q in range(P):
            qy,qx=divmod(q,GW); s=np.zeros(32)
            if fi>=1: s+=gd[fi-1,q].sum(0)
            if fi+1<B:
                for k in range(64):
                    py=qy-((k>>3)-4); px=qx-((k&7)-4)
                    if 0<=py<GH and 0<=px<GW: s-=gd[fi,py*GW+px,k]
            df[fi,q]=s
    assert np.abs(df.reshape(B,GH,GW,32)-f.grad.numpy()).max()<1e-10
    # l2norm backward
    x=torch.tensor(rng.normal(size=(7,32)),dtype=D,requires_grad=True); y=O.l2_normalize(x); gy=torch.tensor(rng.normal(size=(7,32)),dtype=D)
    (y*gy).sum().backward()
    xn=x.detach().numpy(); inv=1/np.sqrt(np.maximum((xn**2).sum(-1,keepdims=True),1e-12)); dot=(xn*inv*gy.numpy()).sum(-1,keepdims=True)
    dx=inv*(gy.numpy()-xn*inv*dot)
    assert np.abs(dx-x.grad.numpy()).max()<1e-12
